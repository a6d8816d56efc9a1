/* Minimal <rpc/types.h> stand-in so the reference `so` sources compile in an
 * image without libtirpc.  TEST INFRASTRUCTURE ONLY (used by oracle/Makefile to
 * build oracle/_ref from /root/reference); never linked into the product. */
#ifndef SO_SHIM_RPC_TYPES_H
#define SO_SHIM_RPC_TYPES_H
typedef int bool_t;
typedef unsigned int u_int;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#endif
