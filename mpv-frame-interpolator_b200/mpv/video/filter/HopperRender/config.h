/* Compile-time configuration of the HopperRender filter — same knobs and defaults as the
 * reference's video/filter/HopperRender/config.h:1-17. */
#ifndef HOPPERRENDER_CONFIG_H
#define HOPPERRENDER_CONFIG_H

/* Quality */
#define MAX_CALC_RES 270        /* flow lattice height limit; fixed in the CUDA library as well  */
#define NUM_ITERATIONS 0        /* 0 = as many window halvings as possible (the only mode built) */
#define MIN_SEARCH_RADIUS 5
#define MAX_SEARCH_RADIUS 16

/* Performance */
#ifndef AUTO_SEARCH_RADIUS_ADJUST   /* -DAUTO_SEARCH_RADIUS_ADJUST=0 pins the radius (reproducible runs) */
#define AUTO_SEARCH_RADIUS_ADJUST 1
#endif
#define UPPER_PERF_BUFFER 1.4
#define LOWER_PERF_BUFFER 1.6

/* Debugging */
#define INC_APP_IND 0           /* the GTK applet is out of scope for the CUDA build             */
#define SAVE_STATS 0

#endif
