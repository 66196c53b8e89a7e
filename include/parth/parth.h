/*
 * include/parth/parth.h -- public umbrella header of the B200 drop-in, same path and role as the
 * reference's include/parth/parth.h:25 (which only includes ParthAPI.h).
 */
#ifndef PARTH_B200_PARTH_H
#define PARTH_B200_PARTH_H
#include "ParthAPI.h"
#endif
