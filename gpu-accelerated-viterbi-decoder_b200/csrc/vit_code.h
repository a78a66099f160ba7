// vit_code.h -- code parameters of this build of the library (reference viterbi.h:61-63: constLen 7, polyn1 0171,
// polyn2 0133).  Shared by the decode kernels, the device-side channel source and the C ABI (vit_code_parameters).
//
// The generator polynomials are compile-time parameters: -DVIT_POLY1=<octal> -DVIT_POLY2=<octal> (csrc/Makefile: EXTRA=)
// builds a decoder and a device source for another K=7 rate-1/2 code.  Like the reference's cores, the stage code relies
// on both polynomials tapping the newest and the oldest encoder bit (bits 6 and 0), which makes the two branches into
// a state carry complementary symbols.  The constraint length is not a parameter: 64 states are the 6-bit position
// space of the kernel's state map.
#pragma once
#if !defined(VIT_POLY1)
#define VIT_POLY1 0171
#endif
#if !defined(VIT_POLY2)
#define VIT_POLY2 0133
#endif
#define VIT_CONST_LEN 7
#if defined(__cplusplus)
static_assert((VIT_POLY1) > 0 && (VIT_POLY1) < 128 && (VIT_POLY2) > 0 && (VIT_POLY2) < 128, "K=7 generator polynomials are 7-bit values");
static_assert(((VIT_POLY1) & 0101) == 0101 && ((VIT_POLY2) & 0101) == 0101,
              "both generator polynomials must tap encoder bits 0 and 6 (complementary branch symbols)");
#endif
