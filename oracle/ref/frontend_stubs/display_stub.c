/* display_stub.c - the window of the headless build: the interactive mode exits at once */
#include "display.h"

void display_create(Display** display, uint32_t width, uint32_t height, bool sync_render_resolution) {
  (void) sync_render_resolution;
  static Display d;
  d.width = width, d.height = height, d.camera_handler = 0;
  *display = &d;
}
void display_query_events(Display* display, DisplayFileDrop** file_drop_array, bool* exit_requested, bool* dirty) {
  (void) display, (void) file_drop_array;
  *exit_requested = true;
  *dirty          = false;
}
void display_handle_inputs(Display* display, LuminaryHost* host, float time_step) { (void) display, (void) host, (void) time_step; }
void display_handle_outputs(Display* display, LuminaryHost* host, const char* output_directory) { (void) display, (void) host, (void) output_directory; }
void display_render(Display* display, LuminaryHost* host) { (void) display, (void) host; }
void display_update(Display* display) { (void) display; }
void display_destroy(Display** display) { *display = 0; }
void camera_handler_center_instance(CameraHandler* camera_handler, LuminaryHost* host, const LuminaryInstance* instance) {
  (void) camera_handler, (void) host, (void) instance;
}
