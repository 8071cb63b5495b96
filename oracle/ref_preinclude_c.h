/* TEST INFRASTRUCTURE (oracle) -- force-included in front of the reference's DEGENSAC C files.
 * exp_ranH.c:823 / exp_ranF.c:822 seed libc with srand(time(NULL)); to get a repeatable
 * hypothesis sequence out of the UNMODIFIED source we route that one call to a settable seed. */
#ifndef MB2_REF_PREINCLUDE_C_H
#define MB2_REF_PREINCLUDE_C_H
#include <time.h>
extern long mb2_ref_seed;
static inline time_t mb2_ref_time(void) { return (time_t)mb2_ref_seed; }
#define time(x) mb2_ref_time()
#endif
