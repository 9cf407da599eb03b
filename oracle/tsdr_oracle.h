/*
 * tsdr_oracle.h -- flat C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The same set of functions exists twice:
 *
 *   orc_*   (oracle/tsdr_oracle.c -> oracle/libtsdr_oracle.so)  my own plain-C restatement of the
 *           reference algorithm, each function citing the reference file:line it follows;
 *   refh_*  (oracle/ref_harness.c -> oracle/_ref/libtsdr_refharness.so)  the REAL reference code, compiled
 *           in place from /root/reference, reached through thin adapters.
 *
 * tests/test_oracle_pinning.py runs both on the same seeded inputs and demands bit-identical results, and
 * tests/golden/ holds outputs of refh_* so the restatement stays pinned on machines without /root/reference.
 *
 * Nothing in the product (tempestsdr_b200/, include/) may include or link this.
 */
#ifndef TSDR_ORACLE_H_
#define TSDR_ORACLE_H_

#include <stdint.h>

#ifndef ORC_PREFIX
#define ORC_PREFIX orc_
#endif
#define ORC_CAT2_(a, b) a##b
#define ORC_CAT_(a, b) ORC_CAT2_(a, b)
#define ORC(name) ORC_CAT_(ORC_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plain-data state blocks (layout shared by both implementations and by ctypes) ---- */

typedef struct { float lastmax, lastmin, snr; } orc_autogain_t;          /* dsp.h:30-34 */
typedef struct { int dx, vx, absvx, curr_stripsize; } orc_sweetspot_t;   /* syncdetector.h:16-22 */
typedef struct { int64_t difference; } orc_dropcomp_t;                   /* dsp.h:91-93 */

typedef struct {
	/* geometry / flags the frame stage reads from tsdr_lib_t */
	uint32_t samplerate;
	int      height;
	double   refreshrate;
	int      autoshift;      /* PARAM_INT_AUTOSHIFT */
	int      pll;            /* PARAM_INT_FRAMERATE_PLL */
	int      superres;       /* PARAM_AUTOCORR_SUPERRESOLUTION */
} orc_pp_config_t;

typedef struct {
	orc_sweetspot_t x, y;
	double  avg_speed;
	int     pll_state;
	float   lastmax, lastmin, snr;
	double  refreshrate_after;   /* tsdr->refreshrate after frameratepll */
	int     width_after;         /* tsdr->width after frameratepll */
	int     pll_callback_fired;  /* VALUE_ID_PLL_FRAMERATE announced this frame */
	int     autogain_callback_fired;
	double  autogain_cb_min, autogain_cb_max;
} orc_pp_result_t;

/* ---- a2: TSDRLibrary.c:244-262 ---- */
void ORC(am_demod)(const float *iq, int pairs, float *out);

/* ---- a6: dsp.c:250-307.  Handle keeps {contrib, offset} and the persistent output buffer ---- */
void    *ORC(resample_new)(void);
void     ORC(resample_free)(void *h);
void     ORC(resample_get)(void *h, double *contrib, double *offset);
void     ORC(resample_set)(void *h, double contrib, double offset);
/* returns output_samples; copies min(output_samples,out_cap) floats to out */
uint32_t ORC(resample_run)(void *h, const float *in, uint32_t size, double upsample_by, double downsample_by,
                           int nearest, float *out, uint32_t out_cap);

#ifndef ORC_IS_HARNESS
/* port only: how many pixels the loop wrote in the last run.  When this is output_samples-1 the reference
 * leaves the last slot stale (zero at first, whatever realloc returns once the buffer has grown). */
uint32_t ORC(resample_last_emitted)(void *h);
#endif

/* ---- a3: dsp.c:313-368 ---- */
void     ORC(dropcomp_shift_with)(orc_dropcomp_t *s, uint32_t block, int64_t syncoffset);
int      ORC(dropcomp_will_drop_all)(orc_dropcomp_t *s, uint32_t size, uint32_t block);
/* ring_accepts: whether cb_add succeeds.  Returns elements forwarded, *skip = leading elements discarded */
uint32_t ORC(dropcomp_add)(orc_dropcomp_t *s, uint32_t size, uint32_t block, int ring_accepts, uint32_t *skip);

/* ---- a4: TSDRLibrary.c:540-550 ---- */
void ORC(geometry)(uint32_t samplerate, int height, double refreshrate, int *width, double *pixelrate,
                   double *pixeltimeoversampletime);

/* ---- a8, a9, a10: dsp.c:22-110 ---- */
void ORC(autogain)(orc_autogain_t *s, int n, const float *in, float *out, float norm);
void ORC(timelowpass)(float coeff, int n, const float *in, float *screen);
void ORC(average_v_h)(int w, int h, const float *in, float *wbuf, float *hbuf);

/* ---- a12-a14: gaussian.c:18-79, syncdetector.c:26-119 ---- */
void ORC(gaussianblur)(float *data, int n);
void ORC(findbestfit)(const float *data, int size, float totalsum, int stripsize, double *bestfit, int *bestfitid);
void ORC(findthesweetspot)(orc_sweetspot_t *s, float *data, int size, int minsize, double lowpasscoeff);

/* ---- a7, a11, a15: dsp.c:134-239, syncdetector.c:133-226 ---- */
void *ORC(pp_new)(void);
void  ORC(pp_free)(void *h);
void  ORC(pp_config)(void *h, const orc_pp_config_t *cfg);
/* frame_in is not modified.  Returns 0 on success. */
int   ORC(pp_run)(void *h, const float *frame_in, int w, int hgt, float motionblur, float lowpasscoeff,
                  int lowpass_before_sync, int autogain_after_proc, float *frame_out, orc_pp_result_t *res);

/* ---- a16: host pixel rule, JavaGUI/jni/TSDRLibraryNDK.c:222-283 (restated only; no JDK here) ---- */
void ORC(pixels_argb)(const float *frame, int n, int inverted, const int32_t *prev, int32_t *argb);

/* ---- a19, a20: fft.c ---- */
uint32_t ORC(fft_getrealsize)(uint32_t size);
void ORC(fft)(float *iq, uint32_t size, int inverse);
void ORC(autocorrelation)(float *answer, const float *real, uint32_t size);
void ORC(crosscorrelation)(float *a_out, float *b_tmp, uint32_t samples);

/* ---- a18: frameratedetector.c:34-62, 87-126 ---- */
void ORC(accumulate)(double *out, uint64_t calls, const float *in_complex, int startid, int length);
/* one capture through frameratedetector_runontodata (autocorrelate + two accumulates + plot callbacks).
 * Plots are copied out (at most *_cap doubles); returns 0 when the plots were produced. */
void *ORC(frd_new)(void);
void  ORC(frd_free)(void *h);
int   ORC(frd_run)(void *h, uint32_t samplerate, const float *data, int size,
                   double *frame_plot, int frame_cap, int *frame_off, int *frame_len,
                   double *line_plot, int line_cap, int *line_off, int *line_len, uint64_t *calls);
#ifndef ORC_IS_HARNESS   /* port-only helpers (pinned through frd_run's offsets/lengths) */
void ORC(framerate_windows)(uint32_t samplerate, int *frame_min, int *frame_max, int *line_min, int *line_max);
uint32_t ORC(framerate_capture_size)(uint32_t samplerate);
#endif

/* ---- a22: superbandwidth.c:67-152 ---- */
void ORC(complex_to_abs_diff)(float *data, int size);
int  ORC(superb_bestfit)(const float *data1, const float *data2, int size, int samples_in_frame);
/* hops[i] holds count_pairs IQ pairs (modified in place like the reference).  out must hold
 * nhops * fft_getrealsize(count_pairs) * 2 floats.  Returns total complex samples written;
 * best_offsets[i] (i>=1) receives the alignment shift in floats. */
int  ORC(superb_ondataready)(float **hops, int nhops, int count_pairs, int samples_in_frame, float *out,
                             int *best_offsets);

#ifdef __cplusplus
}
#endif
#endif
