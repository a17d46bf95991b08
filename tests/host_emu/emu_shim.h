// emu_shim.h — TEST-ONLY shim that lets balance_robot_b200/csrc/brb_kernels.cu's per-env functions compile as
// plain C++ (g++), so CPU tests can compare the kernel's exact fp32 arithmetic with the fp64 oracle.
#ifndef BRB_EMU_SHIM_H
#define BRB_EMU_SHIM_H
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
// one robot per call on the host: the warp-level votes of the per-lane state machine degenerate
#define __any_sync(mask, pred) (pred)
#define __syncwarp(mask) ((void)0)
#define BRB_CTA_OR(p) (p)
#define BRB_CTA_SYNC() ((void)0)
#endif
