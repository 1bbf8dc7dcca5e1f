/* config.h - what CMake would generate from src/mandarin_duck/config.h.in */
#ifndef CONFIG_H
#define CONFIG_H
#include <stdlib.h>
#define MANDARIN_DUCK_VERSION_DATE "n/a"
#define MANDARIN_DUCK_BRANCH_NAME "reference front end on luminary_b200"
#define MANDARIN_DUCK_VERSION_HASH "n/a"
#define MANDARIN_DUCK_VERSION "headless"
#endif /* CONFIG_H */
