// ORACLE-ONLY (test infrastructure): stand-in for the reference's Raytracer/precomp.h
// (which pulls in <windows.h>, robustwin32io and backslash include paths). It lets the
// reference's translation units compile UNMODIFIED with g++ on Linux. Nothing under
// oracle/ is linked into, imported by, or called from the product library.
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <limits.h>
#include <functional>
#include <immintrin.h>

#define __forceinline inline __attribute__((always_inline))

#ifdef ORACLE_WITH_SDL_DECLS
// Declarations only (SDL_main is dead code in the headless build; --gc-sections drops it).
extern "C" {
#include "SDL.h"
}
#include "microui.h"
#endif

#include "MathLib/my_math.h"
#include "MathLib/simd_math_4x.h"
using namespace math;

// platform.h only defines these under _WIN32 (Raytracer/platform.h:37-68); the
// Interlocked* family returns the NEW value, and so do these.
#define WRITE_BARRIER __asm__ __volatile__("" ::: "memory")
#define READ_BARRIER  __asm__ __volatile__("" ::: "memory")

static inline uint32_t atomic_add(volatile uint32_t* addend, int32_t value) {
    return __atomic_add_fetch(addend, (uint32_t)value, __ATOMIC_SEQ_CST);
}
static inline int32_t atomic_add(volatile int32_t* addend, int32_t value) {
    return __atomic_add_fetch(addend, value, __ATOMIC_SEQ_CST);
}
static inline uint64_t atomic_add(volatile uint64_t* addend, uint64_t value) {
    return __atomic_add_fetch(addend, value, __ATOMIC_SEQ_CST);
}
static inline int64_t atomic_add(volatile int64_t* addend, int64_t value) {
    return __atomic_add_fetch(addend, value, __ATOMIC_SEQ_CST);
}
