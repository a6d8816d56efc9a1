/* Minimal <rpc/xdr.h> stand-in (stdio streams, big-endian int/float/double,
 * xdr_vector) — just the calls kd2.c makes for `-std` tipsy files
 * (/root/reference/kd2.c:32-44, 214-233, 333-417, 1293-1330).
 * TEST INFRASTRUCTURE ONLY: lets oracle/Makefile build the untouched reference
 * sources into oracle/_ref without libtirpc.  Not part of the product. */
#ifndef SO_SHIM_RPC_XDR_H
#define SO_SHIM_RPC_XDR_H
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include "types.h"

enum xdr_op { XDR_ENCODE = 0, XDR_DECODE = 1, XDR_FREE = 2 };
typedef struct { enum xdr_op x_op; FILE *x_fp; } XDR;
typedef bool_t (*xdrproc_t)(XDR *, void *, ...);

static inline void xdrstdio_create(XDR *x, FILE *fp, enum xdr_op op) { x->x_op = op; x->x_fp = fp; }
static inline void xdr_destroy_(XDR *x) { if (x->x_op == XDR_ENCODE) fflush(x->x_fp); }
#define xdr_destroy(x) xdr_destroy_(x)

static inline bool_t xdr_shim_word(XDR *x, uint32_t *w)
{
    unsigned char b[4];
    if (x->x_op == XDR_DECODE) {
        if (fread(b, 1, 4, x->x_fp) != 4) return FALSE;
        *w = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
        return TRUE;
    }
    b[0] = (unsigned char)(*w >> 24); b[1] = (unsigned char)(*w >> 16);
    b[2] = (unsigned char)(*w >> 8);  b[3] = (unsigned char)(*w);
    return fwrite(b, 1, 4, x->x_fp) == 4;
}
static inline bool_t xdr_int(XDR *x, int *v)
{
    uint32_t w = (uint32_t)*v;
    if (!xdr_shim_word(x, &w)) return FALSE;
    *v = (int)w; return TRUE;
}
static inline bool_t xdr_float(XDR *x, float *v)
{
    uint32_t w; memcpy(&w, v, 4);
    if (!xdr_shim_word(x, &w)) return FALSE;
    memcpy(v, &w, 4); return TRUE;
}
static inline bool_t xdr_double(XDR *x, double *v)
{
    uint64_t q; uint32_t hi, lo; memcpy(&q, v, 8);
    hi = (uint32_t)(q >> 32); lo = (uint32_t)q;
    if (!xdr_shim_word(x, &hi)) return FALSE;
    if (!xdr_shim_word(x, &lo)) return FALSE;
    q = ((uint64_t)hi << 32) | lo; memcpy(v, &q, 8); return TRUE;
}
static inline bool_t xdr_vector(XDR *x, char *base, u_int n, u_int size, xdrproc_t proc)
{
    u_int i;
    for (i = 0; i < n; ++i) if (!proc(x, base + (size_t)i * size)) return FALSE;
    return TRUE;
}
#endif
