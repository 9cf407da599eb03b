/*
 * TSDRPluginX.h -- OPTIONAL extension beside the unchanged ten-symbol plugin ABI (TSDRPlugin.h): a "raw sink".
 *
 * a front end whose native sample format is 8- or 16-bit integers (TSDRPlugin_RawFile.c:241-261
 * converts them to float on the host before calling back) can hand the block over in its wire format, so that it
 * crosses PCIe at 2 or 4 bytes per IQ pair instead of 8 and is converted on the GPU to exactly the same floats.
 *
 * How it stays a drop-in both ways:
 *   - a plugin MAY export one extra symbol, tsdrpluginx_set_raw_sink.  This library looks it up after the ten
 *     mandatory symbols (TSDRPluginLoader.c:33-72 is otherwise mirrored unchanged) and, if present, calls it once
 *     before tsdrplugin_init with the table below.  The reference library never calls it: the plugin then converts
 *     on the host and uses the ordinary float callback, i.e. it is an ordinary TSDR plugin.
 *   - a plugin that does not export it is driven exactly as before.
 */
#ifndef TSDR_PLUGIN_X_H
#define TSDR_PLUGIN_X_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sample formats: the RawFile plugin's own numbering (TSDRPlugin_RawFile.c:29-33) */
#define TSDRX_FMT_FLOAT  0
#define TSDRX_FMT_INT8   1
#define TSDRX_FMT_INT16  2
#define TSDRX_FMT_UINT8  3
#define TSDRX_FMT_UINT16 4

typedef struct tsdrx_raw_sink {
	uint32_t abi_version;                      /* 1 */
	/* Same meaning as the float callback (TSDRPlugin.h): items_count interleaved I,Q components (even), ctx = the ctx
	 * given to tsdrplugin_readasync, samples_dropped = IQ pairs lost before this block.  The buffer may be reused as
	 * soon as the call returns.  Returns 0, or non-zero when the library has stopped accepting data. */
	int   (*ingest)(const void *samples, int fmt, uint64_t items_count, void *ctx, int64_t samples_dropped);
	/* page-locked host memory for the plugin's read buffer (faster DMA), valid inside tsdrplugin_readasync with its
	 * ctx; either may be NULL or fail: use malloc then */
	void *(*alloc_host)(size_t bytes, void *ctx);
	void  (*free_host)(void *p, void *ctx);
} tsdrx_raw_sink_t;

typedef void (*tsdrpluginx_set_raw_sink_fn)(const tsdrx_raw_sink_t *sink);   /* sink == NULL: forget it */

#ifdef __cplusplus
}
#endif
#endif
