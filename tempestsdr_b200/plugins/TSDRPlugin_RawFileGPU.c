/*
 * TSDRPlugin_RawFileGPU -- a file-playback front end for the TSDR plugin ABI that can hand its samples over in the
 * file's own format.
 *
 * It exports the ten symbols every TSDR plugin exports (include/TSDRPlugin.h) and takes the same parameter string as
 * the reference's TSDRPlugin_RawFile ("filename samplerate format", format = float|int8|uint8|int16|uint16;
 * TSDRPlugin_RawFile.c:163-199), plus two optional words: "nopace" (do not sleep to real time -- the reference has
 * this as a compile-time switch, TSDRPlugin_RawFile.c:35) and "block=<items>" (items per callback, default 524288 as
 * TSDRPlugin_RawFile.c:39).  Loaded by the reference library it behaves like TSDRPlugin_RawFile: blocks are converted
 * to float on the host with the same expressions (TSDRPlugin_RawFile.c:241-261) and delivered through the float
 * callback, and at the end of the file it rewinds and delivers the block buffer as it stands, like the reference loop.
 * Loaded by this repository's library it additionally receives a raw sink (include/TSDRPluginX.h) and passes 8/16-bit
 * blocks through untouched; the GPU produces the same floats.
 *
 * Written from the ABI and the behaviour described above; no code is shared with the reference plugin.
 */
#include "TSDRPlugin.h"
#include "TSDRPluginX.h"
#include "TSDRCodes.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define DEFAULT_BLOCK_ITEMS (512u * 1024u)
#define MAX_RATE 1000e6

static struct {
	char path[1024];
	uint32_t rate;
	int fmt, bytes_per_item;
	int paced;
	uint64_t block_items;
	volatile int running;
	const tsdrx_raw_sink_t *sink;
	char *errtext; int errcode;
} S = { .fmt = -1, .paced = 1, .block_items = DEFAULT_BLOCK_ITEMS };

static int set_error(int code, const char *text) {
	S.errcode = code;
	free(S.errtext);
	S.errtext = (code == TSDR_OK || !text) ? NULL : strdup(text);
	return code;
}

/* ---- parameter string: words separated by blanks; '...' or "..." keep blanks inside a word ----------------------- */
static int next_word(const char **cursor, char *out, size_t cap) {
	const char *p = *cursor;
	while (*p == ' ') p++;
	if (!*p) { *cursor = p; return 0; }
	size_t n = 0;
	char quote = 0;
	for (; *p; p++) {
		if (quote) { if (*p == quote) { quote = 0; continue; } }
		else if (*p == '\'' || *p == '"') { quote = *p; continue; }
		else if (*p == ' ') break;
		if (n + 1 < cap) out[n++] = *p;
	}
	out[n] = 0;
	*cursor = p;
	return quote ? -1 : 1;                       /* -1: unterminated quote */
}

TSDRPLUGIN_API void __stdcall tsdrplugin_getName(char *name) { strcpy(name, "TSDR Raw File (GPU-aware) Plugin"); }

TSDRPLUGIN_API void tsdrpluginx_set_raw_sink(const tsdrx_raw_sink_t *sink) {
	S.sink = (sink && sink->abi_version >= 1 && sink->ingest) ? sink : NULL;
}

TSDRPLUGIN_API int __stdcall tsdrplugin_init(const char *params) {
	static const char *usage = "Parameters: filename samplerate format [nopace] [block=items]; format is float, int8, uint8, int16 or uint16.";
	char word[1024];
	const char *cur = params ? params : "";
	if (next_word(&cur, word, sizeof word) <= 0) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, usage);
	snprintf(S.path, sizeof S.path, "%s", word);
	if (next_word(&cur, word, sizeof word) <= 0) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, usage);
	const long rate = atol(word);
	if (rate <= 0 || (double) rate > MAX_RATE) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, "The sample rate of the recording is invalid.");
	if (next_word(&cur, word, sizeof word) <= 0) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, usage);
	static const struct { const char *name; int fmt, bytes; } formats[] = {
		{"float", TSDRX_FMT_FLOAT, 4}, {"int8", TSDRX_FMT_INT8, 1}, {"int16", TSDRX_FMT_INT16, 2},
		{"uint8", TSDRX_FMT_UINT8, 1}, {"uint16", TSDRX_FMT_UINT16, 2},
	};
	S.fmt = -1;
	for (size_t i = 0; i < sizeof formats / sizeof formats[0]; i++)
		if (!strcmp(word, formats[i].name)) { S.fmt = formats[i].fmt; S.bytes_per_item = formats[i].bytes; }
	if (S.fmt < 0) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, usage);
	S.paced = 1; S.block_items = DEFAULT_BLOCK_ITEMS;
	int got;
	while ((got = next_word(&cur, word, sizeof word)) > 0) {
		if (!strcmp(word, "nopace")) S.paced = 0;
		else if (!strncmp(word, "block=", 6)) {
			const long long b = atoll(word + 6);
			if (b < 2 || (b & 1) || b > (1ll << 30)) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, "block= must be an even number of items between 2 and 2^30.");
			S.block_items = (uint64_t) b;
		} else return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, usage);
	}
	if (got < 0) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, usage);
	S.rate = (uint32_t) rate;
	return set_error(TSDR_OK, NULL);
}

TSDRPLUGIN_API uint32_t __stdcall tsdrplugin_setsamplerate(uint32_t rate) { (void) rate; return S.rate; }   /* a recording has one rate */
TSDRPLUGIN_API uint32_t __stdcall tsdrplugin_getsamplerate(void) { return S.rate; }
TSDRPLUGIN_API int __stdcall tsdrplugin_setbasefreq(uint32_t freq) { (void) freq; return set_error(TSDR_OK, NULL); }
TSDRPLUGIN_API int __stdcall tsdrplugin_setgain(float gain) { (void) gain; return set_error(TSDR_OK, NULL); }
TSDRPLUGIN_API int __stdcall tsdrplugin_stop(void) { S.running = 0; return set_error(TSDR_OK, NULL); }
TSDRPLUGIN_API char *__stdcall tsdrplugin_getlasterrortext(void) { return S.errcode == TSDR_OK ? NULL : S.errtext; }
TSDRPLUGIN_API void __stdcall tsdrplugin_cleanup(void) { S.sink = NULL; }

/* host conversion, used only when no raw sink was offered (TSDRPlugin_RawFile.c:241-261) */
static void to_float(const void *raw, float *out, uint64_t items, int fmt) {
	switch (fmt) {
	case TSDRX_FMT_INT8:   { const int8_t *p = raw;   for (uint64_t i = 0; i < items; i++) out[i] = (float) (p[i] / 128.0); } break;
	case TSDRX_FMT_UINT8:  { const uint8_t *p = raw;  for (uint64_t i = 0; i < items; i++) out[i] = (float) ((p[i] - 128) / 128.0); } break;
	case TSDRX_FMT_INT16:  { const int16_t *p = raw;  for (uint64_t i = 0; i < items; i++) out[i] = (float) (p[i] / 32767.0); } break;
	case TSDRX_FMT_UINT16: { const uint16_t *p = raw; for (uint64_t i = 0; i < items; i++) out[i] = (float) ((p[i] - 32767) / 32767.0); } break;
	default: memcpy(out, raw, sizeof(float) * items);
	}
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

TSDRPLUGIN_API int __stdcall tsdrplugin_readasync(tsdrplugin_readasync_function cb, void *ctx) {
	if (S.fmt < 0) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, "The plugin has not been initialised.");
	FILE *f = fopen(S.path, "rb");
	if (!f) return set_error(TSDR_PLUGIN_PARAMETERS_WRONG, "Cannot open the recording.");
	const tsdrx_raw_sink_t *sink = S.sink;
	const size_t block_bytes = (size_t) S.block_items * (size_t) S.bytes_per_item;
	int pinned = 0;
	unsigned char *raw = NULL;
	if (sink && sink->alloc_host && sink->free_host && (raw = sink->alloc_host(block_bytes, ctx)) != NULL) pinned = 1;
	if (!raw) raw = malloc(block_bytes);
	float *conv = (!sink && S.fmt != TSDRX_FMT_FLOAT) ? malloc(sizeof(float) * S.block_items) : NULL;
	if (!raw || (!sink && S.fmt != TSDRX_FMT_FLOAT && !conv)) {
		if (raw) { if (pinned) sink->free_host(raw, ctx); else free(raw); }
		free(conv); fclose(f);
		return set_error(TSDR_ERR_PLUGIN, "Out of memory.");
	}
	memset(raw, 0, block_bytes);
	const double block_seconds = (double) S.block_items / (double) S.rate;     /* real-time pacing, as TSDRPlugin_RawFile.c:221-222 */
	S.running = 1;
	double due = now_s();
	while (S.running) {
		/* fill the block; at the end of the file rewind and deliver the buffer as it stands */
		size_t have = 0;
		while (have < block_bytes) {
			const size_t got = fread(raw + have, 1, block_bytes - have, f);
			have += got;
			if (got == 0) { rewind(f); break; }
		}
		if (!S.running) break;
		if (sink) {
			if (sink->ingest(raw, S.fmt, S.block_items, ctx, 0) != 0) break;
		} else if (S.fmt == TSDRX_FMT_FLOAT) cb((float *) raw, S.block_items, ctx, 0);
		else { to_float(raw, conv, S.block_items, S.fmt); cb(conv, S.block_items, ctx, 0); }
		if (S.paced) {
			due += block_seconds;
			const double wait = due - now_s();
			if (wait > 0) { struct timespec ts = { (time_t) wait, (long) ((wait - (double) (time_t) wait) * 1e9) }; nanosleep(&ts, NULL); }
			else due = now_s();
		}
	}
	if (pinned) sink->free_host(raw, ctx); else free(raw);
	free(conv);
	fclose(f);
	S.running = 0;
	return set_error(TSDR_OK, NULL);
}
