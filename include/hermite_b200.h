/* hermite_b200.h -- C ABI of libhermite_b200.so (work in progress).
 *
 * Drop-in boundary for the numerical core of urbainvaes/hermipy: every entry point replaces one
 * function that the reference registers in its Boost.Python module `hermite_cpp`
 * (cpp/src/hermite/python/module.cpp:45-168) and that hermipy/core.py calls.
 * Plain pointers and sizes only; every function returns an int status and never calls exit().
 */
#ifndef HERMITE_B200_H
#define HERMITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define HB_OK 0
#define HB_ERR_INVALID 1   /* bad argument (reference: message + exit(1)) */
#define HB_ERR_INDEX_SET 2 /* unknown index set (reference: "Invalid index set!" + exit(1)) */
#define HB_ERR_FOURIER_DEGREE 3 /* odd degree on a Fourier axis (transform.cpp:92-96) */
#define HB_ERR_DEGREE 4    /* n_polys does not match any degree (lib.cpp:93-99: exit(0)) */
#define HB_ERR_CUDA 5      /* CUDA runtime/driver failure; hb_last_error() has the text */
#define HB_ERR_ALIGN 6     /* device operand violates the 16-byte TMA alignment rule */
#define HB_ERR_NOMEM 7

#ifdef __cplusplus
}
#endif
#endif
