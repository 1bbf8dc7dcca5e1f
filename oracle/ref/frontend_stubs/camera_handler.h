/* camera_handler.h - stand-in, see display.h */
#ifndef MANDARIN_DUCK_CAMERA_HANDLER_H
#define MANDARIN_DUCK_CAMERA_HANDLER_H

#include "utils.h"

typedef struct CameraHandler CameraHandler;
void camera_handler_center_instance(CameraHandler* camera_handler, LuminaryHost* host, const LuminaryInstance* instance);

#endif /* MANDARIN_DUCK_CAMERA_HANDLER_H */
