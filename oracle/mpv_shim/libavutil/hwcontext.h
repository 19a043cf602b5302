/* shim: see ../hr_mpv_shim.h */
#include "../hr_mpv_shim.h"
