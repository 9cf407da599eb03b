/*
 * TSDRPlugin.h -- the source-plugin ABI (what an SDR front-end plugin exports and what the library dlsym()s).
 *
 * Contract identical to the reference (TempestSDR/src/include/TSDRPlugin.h:49-60; loader TSDRPluginLoader.c:33-72):
 * ten C symbols, cdecl on Linux.  Existing plugins (TSDRPlugin_RawFile, _UHD, _Mirics, _SdrPlay, _ExtIO) load into
 * this library unchanged; nothing here is specific to the GPU implementation.
 *
 * The data callback: `buf` holds items_count floats = items_count/2 interleaved I,Q pairs (items_count even, may be
 * 0), owned by the plugin, valid only during the call and allowed to be modified by the callee;
 * samples_dropped = IQ pairs lost before this block.
 */
#ifndef _TSDRPluginHeader
#define _TSDRPluginHeader

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32) || defined(__CYGWIN__)
  #define TSDRPLUGIN_API __declspec(dllexport)
#elif defined(__GNUC__) && __GNUC__ >= 4
  #define TSDRPLUGIN_API __attribute__((visibility("default")))
#else
  #define TSDRPLUGIN_API
#endif
#if !(defined(_WIN32) || defined(_WIN64) || defined(__stdcall))
  #define __stdcall
#endif

typedef void (*tsdrplugin_readasync_function)(float *buf, uint64_t items_count, void *ctx, int64_t samples_dropped);

TSDRPLUGIN_API void     __stdcall tsdrplugin_getName(char *name);
TSDRPLUGIN_API int      __stdcall tsdrplugin_init(const char *params);
TSDRPLUGIN_API uint32_t __stdcall tsdrplugin_setsamplerate(uint32_t rate);
TSDRPLUGIN_API uint32_t __stdcall tsdrplugin_getsamplerate(void);
TSDRPLUGIN_API int      __stdcall tsdrplugin_setbasefreq(uint32_t freq);
TSDRPLUGIN_API int      __stdcall tsdrplugin_stop(void);
TSDRPLUGIN_API int      __stdcall tsdrplugin_setgain(float gain);
TSDRPLUGIN_API char *   __stdcall tsdrplugin_getlasterrortext(void);
TSDRPLUGIN_API int      __stdcall tsdrplugin_readasync(tsdrplugin_readasync_function cb, void *ctx);
TSDRPLUGIN_API void     __stdcall tsdrplugin_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
