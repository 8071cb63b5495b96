/* TEST INFRASTRUCTURE (oracle) -- force-included in front of the reference's DEGENSAC C files.
 * exp_ranH.c:823 / exp_ranF.c:822 seed libc with srand(time(NULL)); to get a repeatable
 * hypothesis sequence out of the UNMODIFIED source we route that one call to a settable seed. */
#ifndef MB2_REF_PREINCLUDE_C_H
#define MB2_REF_PREINCLUDE_C_H
#include <time.h>
extern long mb2_ref_seed;
static inline time_t mb2_ref_time(void) { return (time_t)mb2_ref_seed; }
#define time(x) mb2_ref_time()
/* degensac/lapwrap.h:12 declares `typedef ptrdiff_t lapack_int` and passes pointers to such 64-bit integers to an LP64 LAPACK
 * (32-bit INTEGER): sizes work by little-endian accident, but `info` keeps an UNINITIALISED upper half, so lap_SVD / lap_eig
 * report failure depending on stack garbage (singulF then silently returns the identity, Ftools.c:294-297).  The oracle build
 * pins the evidently intended behaviour -- LAPACK's own 32-bit info decides -- by making lapack_int a 32-bit int. */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#define ptrdiff_t int
#endif
