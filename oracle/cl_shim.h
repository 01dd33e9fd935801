/* cl_shim.h — lets gcc compile the reference's OpenCL C kernel files AS C, in place, so that the
 * reference's own kernel code can be executed on the CPU (TEST INFRASTRUCTURE ONLY; see Makefile
 * target _ref/libref.so).  One work-item runs at a time: get_global_id() etc. return what the
 * driver (ref_driver.c) has set, barriers are no-ops (a single work-item needs none), __local
 * variables are statics of the one "work-group", atomics are plain read-modify-writes.
 * Launching process_coordinates with ONE work-item (global size 1) is a legal NDRange for that
 * kernel (its loop strides by get_global_size) and gives the sequential emission order the
 * contract calls canonical. */
#ifndef EVK_CL_SHIM_H_
#define EVK_CL_SHIM_H_
#include <math.h>
#include <stdatomic.h>
#include <stddef.h>

#define __kernel
#define __global
#ifndef __local
#define __local static /* function-scope work-group variables; -D__local= for pointer qualifiers */
#endif
#define __constant const
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2

typedef unsigned int uint;
typedef unsigned char uchar;
typedef struct { float x, y, z; } float3;

extern _Thread_local size_t ref_gid, ref_gsize, ref_lid, ref_lsize, ref_group;
static inline size_t get_global_id(int d) { (void)d; return ref_gid; }
static inline size_t get_global_size(int d) { (void)d; return ref_gsize; }
static inline size_t get_local_id(int d) { (void)d; return ref_lid; }
static inline size_t get_local_size(int d) { (void)d; return ref_lsize; }
static inline size_t get_group_id(int d) { (void)d; return ref_group; }
static inline void barrier(int flags) { (void)flags; }

static inline int atomic_inc(volatile int* p) { int old = *p; *p = old + 1; return old; }
static inline int atomic_add(volatile int* p, int v) { int old = *p; *p = old + v; return old; }

static inline float3 make_float3(float x, float y, float z) { float3 v = {x, y, z}; return v; }
/* OpenCL length(): sqrt of the sum of squares, single precision */
static inline float length(float3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
#endif
