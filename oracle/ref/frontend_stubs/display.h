/* display.h - stand-in for the reference front end's SDL window (src/mandarin_duck/display.h) when its command line / benchmark half
 * (main.c, argument_parser.c, mandarin_duck.c) is built headless against include/luminary by oracle/ref/Makefile: frontend.
 * Only what mandarin_duck.c names; the interactive mode is not reachable in the tests. Test infrastructure, never shipped. */
#ifndef MANDARIN_DUCK_DISPLAY_H
#define MANDARIN_DUCK_DISPLAY_H

#include "camera_handler.h"
#include "utils.h"

typedef struct DisplayFileDrop {
  const char* file_path;
} DisplayFileDrop;

typedef struct Display {
  uint32_t width;
  uint32_t height;
  CameraHandler* camera_handler;
} Display;

void display_create(Display** display, uint32_t width, uint32_t height, bool sync_render_resolution);
void display_query_events(Display* display, DisplayFileDrop** file_drop_array, bool* exit_requested, bool* dirty);
void display_handle_inputs(Display* display, LuminaryHost* host, float time_step);
void display_handle_outputs(Display* display, LuminaryHost* host, const char* output_directory);
void display_render(Display* display, LuminaryHost* host);
void display_update(Display* display);
void display_destroy(Display** display);

#endif /* MANDARIN_DUCK_DISPLAY_H */
