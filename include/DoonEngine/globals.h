/* DoonEngine/globals.h -- integer vector types, the message callback and the allocator hooks of the
 * DN_* C API, as exported by libdoon_b200.so.
 *
 * Replaces (layout- and name-compatible with) /root/reference/src/DoonEngine/globals.h:
 *   DNivec2..DNuvec4            globals.h:11-40
 *   DNmessageType / Severity    globals.h:46-60   (enumerator order is ABI: the callback receives the ints)
 *   g_DN_message_callback       globals.h:63      (defined in the library; the application assigns it
 *                                                  BEFORE DN_init, it is called without a NULL check upstream;
 *                                                  this library tolerates NULL and then stays silent)
 *   DN_MALLOC/DN_FREE/DN_REALLOC globals.h:69-81
 */
#ifndef DN_GLOBALS_H
#define DN_GLOBALS_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DNivec2 { int32_t x, y; } DNivec2;
typedef struct DNivec3 { int32_t x, y, z; } DNivec3;
typedef struct DNivec4 { int32_t x, y, z, w; } DNivec4;
typedef struct DNuvec2 { uint32_t x, y; } DNuvec2;
typedef struct DNuvec3 { uint32_t x, y, z; } DNuvec3;
typedef struct DNuvec4 { uint32_t x, y, z, w; } DNuvec4;

/* what a message is about */
typedef enum DNmessageType
{
	DN_MESSAGE_CPU_MEMORY = 0,
	DN_MESSAGE_GPU_MEMORY = 1, /* also carries every CUDA error of this library */
	DN_MESSAGE_SHADER     = 2, /* kept for ABI compatibility; never emitted (kernels are compiled ahead of time) */
	DN_MESSAGE_FILE_IO    = 3
} DNmessageType;

typedef enum DNmessageSeverity
{
	DN_MESSAGE_NOTE  = 0, /* informational, e.g. an automatic capacity doubling */
	DN_MESSAGE_ERROR = 1, /* the call failed, the engine keeps running */
	DN_MESSAGE_FATAL = 2  /* the engine cannot continue */
} DNmessageSeverity;

extern void (*g_DN_message_callback)(DNmessageType, DNmessageSeverity, const char*);

#ifndef DN_MALLOC
#define DN_MALLOC(s) malloc((s))
#endif
#ifndef DN_FREE
#define DN_FREE(p) free((p))
#endif
#ifndef DN_REALLOC
#define DN_REALLOC(p, s) realloc((p), (s))
#endif

#ifdef __cplusplus
}
#endif

#endif
