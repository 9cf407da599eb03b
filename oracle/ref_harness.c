/*
 * ref_harness.c -- adapters from the flat oracle API (tsdr_oracle.h, prefix refh_) onto the REAL reference.
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/libtsdr_refharness.so and only when
 * /root/reference is present.  The four reference translation units that hide `static` functions on the hot
 * path are #included in place (nothing is copied); the rest of the reference is compiled beside this file.
 *
 * Everything here is single-threaded and drives the reference's stage functions synchronously, which is the
 * only way to get timing-independent results out of it.
 */
#define ORC_PREFIX refh_
#define ORC_IS_HARNESS 1

/* the reference, in place (-I/root/reference/TempestSDR/src) */
#include "TSDRLibrary.c"     /* am_demod (static inline), process, set_internal_samplerate, tsdr_* */
#include "dsp.c"             /* dsp_* incl. static dsp_dropped_cal_compensation */
#include "syncdetector.c"    /* findbestfit (static inline), findthesweetspot, syncdetector_run */
#include "superbandwidth.c"  /* superb_bestfit (static inline), superb_ondataready */
#include "fft.h"
#include "gaussian.h"

#include "tsdr_oracle.h"
#include <string.h>

/* ------------------------------------------------------------------ a2 */
void refh_am_demod(const float *iq, int pairs, float *out) {
	float *tmp = (float *) malloc(sizeof(float) * 2 * (pairs > 0 ? pairs : 1));
	memcpy(tmp, iq, sizeof(float) * 2 * pairs);
	am_demod(tmp, pairs);               /* in place, first half */
	memcpy(out, tmp, sizeof(float) * pairs);
	free(tmp);
}

/* ------------------------------------------------------------------ a6 */
typedef struct { dsp_resample_t res; extbuffer_t out; } refh_resampler_t;

void *refh_resample_new(void) {
	refh_resampler_t *r = (refh_resampler_t *) calloc(1, sizeof(refh_resampler_t));
	dsp_resample_init(&r->res);
	extbuffer_init(&r->out);
	return r;
}
void refh_resample_free(void *h) {
	refh_resampler_t *r = (refh_resampler_t *) h;
	extbuffer_free(&r->out);
	free(r);
}
void refh_resample_get(void *h, double *contrib, double *offset) {
	refh_resampler_t *r = (refh_resampler_t *) h;
	*contrib = r->res.contrib; *offset = r->res.offset;
}
void refh_resample_set(void *h, double contrib, double offset) {
	refh_resampler_t *r = (refh_resampler_t *) h;
	r->res.contrib = contrib; r->res.offset = offset;
}
uint32_t refh_resample_run(void *h, const float *in, uint32_t size, double upsample_by, double downsample_by,
                           int nearest, float *out, uint32_t out_cap) {
	refh_resampler_t *r = (refh_resampler_t *) h;
	extbuffer_t inb;
	extbuffer_init(&inb);
	inb.buffer = (float *) in;          /* only read by dsp_resample_process */
	inb.size_valid_elements = size;
	inb.buffer_max_size = size;
	inb.valid = 1;
	/* the reference can emit one pixel more than it sized the buffer for (dsp.c:262 vs :288-297); give the
	 * allocation some slack so that quirk stays a harmless write */
	{
		const double r_ = upsample_by / downsample_by;
		const uint32_t want = (uint32_t) (int) ((size - r->res.offset) * r_);
		if (want == 0) return 0;   /* the reference would abort in extbuffer_preparetohandle's assert(size > 0) */
		if (want > 0 && r->out.buffer_max_size >= want && r->out.buffer_max_size <= (want << 1)) {
			/* no realloc will happen inside; make sure one element of slack exists */
			r->out.buffer = (float *) realloc(r->out.buffer, sizeof(float) * ((size_t) r->out.buffer_max_size + 8));
		}
	}
	dsp_resample_process(&r->res, &inb, &r->out, upsample_by, downsample_by, nearest);
	const uint32_t n = r->out.size_valid_elements;
	memcpy(out, r->out.buffer, sizeof(float) * (n < out_cap ? n : out_cap));
	return n;
}

/* ------------------------------------------------------------------ a3 */
void refh_dropcomp_shift_with(orc_dropcomp_t *s, uint32_t block, int64_t syncoffset) {
	dsp_dropped_compensation_t d; d.difference = s->difference;
	dsp_dropped_compensation_shift_with(&d, block, syncoffset);
	s->difference = d.difference;
}
int refh_dropcomp_will_drop_all(orc_dropcomp_t *s, uint32_t size, uint32_t block) {
	dsp_dropped_compensation_t d; d.difference = s->difference;
	return dsp_dropped_compensation_will_drop_all(&d, size, block);
}
uint32_t refh_dropcomp_add(orc_dropcomp_t *s, uint32_t size, uint32_t block, int ring_accepts, uint32_t *skip) {
	dsp_dropped_compensation_t d; d.difference = s->difference;
	CircBuff_t cb;
	cb_init(&cb, CB_SIZE_MAX_COEFF_LOW_LATENCY);
	if (!ring_accepts) cb.invalid = 1;   /* cb_add -> CB_ERROR, the "ring refused the block" branch */
	float *buf = (float *) malloc(sizeof(float) * (size ? size : 1));
	uint32_t i;
	for (i = 0; i < size; i++) buf[i] = (float) i;
	dsp_dropped_compensation_add(&d, &cb, buf, size, block);
	uint32_t forwarded = 0; *skip = 0;
	if ((int64_t) size <= s->difference) *skip = size;      /* whole block swallowed by the debt */
	else if (ring_accepts) {
		forwarded = (uint32_t) cb_size(&cb);
		if (forwarded) {
			float first;
			cb_rem_nonblocking(&cb, &first, 1);
			*skip = (uint32_t) first;
		}
	}
	if (!ring_accepts) cb.invalid = 0;
	cb_free(&cb);
	free(buf);
	s->difference = d.difference;
	return forwarded;
}

/* ------------------------------------------------------------------ a4 */
static void refh_nullvalue(int id, double a, double b, void *ctx) { (void) id; (void) a; (void) b; (void) ctx; }

void refh_geometry(uint32_t samplerate, int height, double refreshrate, int *width, double *pixelrate,
                   double *pixeltimeoversampletime) {
	tsdr_lib_t *t;
	tsdr_init(&t, refh_nullvalue, NULL, NULL);
	t->pixeltimeoversampletime = 0; t->width = 0; t->pixelrate = 0;
	t->height = height; t->refreshrate = refreshrate;
	set_internal_samplerate(t, samplerate);
	*width = t->width; *pixelrate = t->pixelrate; *pixeltimeoversampletime = t->pixeltimeoversampletime;
	t->errormsg = NULL;
	tsdr_free(&t);
}

/* ------------------------------------------------------------------ a8-a10 */
void refh_autogain(orc_autogain_t *s, int n, const float *in, float *out, float norm) {
	dsp_autogain_t a; a.lastmax = s->lastmax; a.lastmin = s->lastmin; a.snr = s->snr;
	dsp_autogain_run(&a, n, (float *) in, out, norm);
	s->lastmax = a.lastmax; s->lastmin = a.lastmin; s->snr = a.snr;
}
void refh_timelowpass(float coeff, int n, const float *in, float *screen) {
	dsp_timelowpass_run(coeff, n, (float *) in, screen);
}
void refh_average_v_h(int w, int h, const float *in, float *wbuf, float *hbuf) {
	dsp_average_v_h(w, h, (float *) in, wbuf, hbuf);
}

/* ------------------------------------------------------------------ a12-a14 */
void refh_gaussianblur(float *data, int n) { gaussianblur(data, n); }
void refh_findbestfit(const float *data, int size, float totalsum, int stripsize, double *bestfit, int *bestfitid) {
	findbestfit((float *) data, size, totalsum, stripsize, bestfit, bestfitid);
}
void refh_findthesweetspot(orc_sweetspot_t *s, float *data, int size, int minsize, double lowpasscoeff) {
	sweetspot_data_t d; d.dx = s->dx; d.vx = s->vx; d.absvx = s->absvx; d.curr_stripsize = s->curr_stripsize;
	findthesweetspot(&d, data, size, minsize, lowpasscoeff);
	s->dx = d.dx; s->vx = d.vx; s->absvx = d.absvx; s->curr_stripsize = d.curr_stripsize;
}

/* ------------------------------------------------------------------ a7, a11, a15 */
typedef struct {
	tsdr_lib_t *tsdr;
	float *scratch; int scratch_n;
	int pll_fired, ag_fired; double ag_min, ag_max;
} refh_pp_t;

static void refh_pp_value(int id, double a, double b, void *ctx) {
	refh_pp_t *p = (refh_pp_t *) ctx;
	if (id == VALUE_ID_PLL_FRAMERATE) p->pll_fired = 1;
	if (id == VALUE_ID_AUTOGAIN_VALUES) { p->ag_fired = 1; p->ag_min = a; p->ag_max = b; }
}

void *refh_pp_new(void) {
	refh_pp_t *p = (refh_pp_t *) calloc(1, sizeof(refh_pp_t));
	tsdr_init(&p->tsdr, refh_pp_value, NULL, p);
	/* tsdr_init leaves these uninitialised (TSDRLibrary.c:62-94) */
	p->tsdr->motionblur = 0; p->tsdr->gain = 0; p->tsdr->height = 1; p->tsdr->refreshrate = 1;
	p->tsdr->width = 0; p->tsdr->samplerate = 0; p->tsdr->pixeltimeoversampletime = 0; p->tsdr->errormsg = NULL;
	p->tsdr->pixelrate = 0; p->tsdr->samplerate_real = 0;
	return p;
}
void refh_pp_free(void *h) {
	refh_pp_t *p = (refh_pp_t *) h;
	tsdr_free(&p->tsdr);
	free(p->scratch);
	free(p);
}
void refh_pp_config(void *h, const orc_pp_config_t *cfg) {
	refh_pp_t *p = (refh_pp_t *) h;
	p->tsdr->height = cfg->height;
	p->tsdr->refreshrate = cfg->refreshrate;
	p->tsdr->samplerate_real = cfg->samplerate;
	set_internal_samplerate(p->tsdr, cfg->samplerate);
	p->tsdr->params_int[PARAM_INT_AUTOSHIFT] = cfg->autoshift;
	p->tsdr->params_int[PARAM_INT_FRAMERATE_PLL] = cfg->pll;
	p->tsdr->params_int[PARAM_AUTOCORR_SUPERRESOLUTION] = cfg->superres;
}
int refh_pp_run(void *h, const float *frame_in, int w, int hgt, float motionblur, float lowpasscoeff,
                int lowpass_before_sync, int autogain_after_proc, float *frame_out, orc_pp_result_t *res) {
	refh_pp_t *p = (refh_pp_t *) h;
	const int n = w * hgt;
	if (n > p->scratch_n) { p->scratch = (float *) realloc(p->scratch, sizeof(float) * n); p->scratch_n = n; }
	memcpy(p->scratch, frame_in, sizeof(float) * n);
	p->pll_fired = 0; p->ag_fired = 0; p->ag_min = 0; p->ag_max = 0;
	dsp_postprocess_t *pp = &p->tsdr->dsp_postprocess;
	float *result = dsp_post_process(p->tsdr, pp, p->scratch, w, hgt, motionblur, lowpasscoeff,
	                                 lowpass_before_sync, autogain_after_proc);
	memcpy(frame_out, result, sizeof(float) * n);
	if (res) {
		res->x.dx = pp->sync.db_x.dx; res->x.vx = pp->sync.db_x.vx; res->x.absvx = pp->sync.db_x.absvx;
		res->x.curr_stripsize = pp->sync.db_x.curr_stripsize;
		res->y.dx = pp->sync.db_y.dx; res->y.vx = pp->sync.db_y.vx; res->y.absvx = pp->sync.db_y.absvx;
		res->y.curr_stripsize = pp->sync.db_y.curr_stripsize;
		res->avg_speed = pp->sync.avg_speed; res->pll_state = pp->sync.state;
		res->lastmax = pp->dsp_autogain.lastmax; res->lastmin = pp->dsp_autogain.lastmin;
		res->snr = pp->dsp_autogain.snr;
		res->refreshrate_after = p->tsdr->refreshrate; res->width_after = p->tsdr->width;
		res->pll_callback_fired = p->pll_fired; res->autogain_callback_fired = p->ag_fired;
		res->autogain_cb_min = p->ag_min; res->autogain_cb_max = p->ag_max;
	}
	return 0;
}

/* ------------------------------------------------------------------ a19, a20 */
uint32_t refh_fft_getrealsize(uint32_t size) { return fft_getrealsize(size); }
void refh_fft(float *iq, uint32_t size, int inverse) { fft_perform(iq, size, inverse); }
void refh_autocorrelation(float *answer, const float *real, uint32_t size) {
	fft_autocorrelation(answer, (float *) real, size);
}
void refh_crosscorrelation(float *a_out, float *b_tmp, uint32_t samples) {
	fft_crosscorrelation(a_out, b_tmp, samples);
}

/* ------------------------------------------------------------------ a18 */
void accummulate(extbuffer_t *out, extbuffer_t *in, int startid, int length);   /* frameratedetector.c:34 */
void frameratedetector_runontodata(frameratedetector_t *frameratedetector, float *data, int size,
                                   extbuffer_t *extbuff, extbuffer_t *extbuff_small1, extbuffer_t *extbuff_small2);

void refh_accumulate(double *out, uint64_t calls, const float *in_complex, int startid, int length) {
	extbuffer_t in, o;
	extbuffer_init(&in);
	in.buffer = (float *) in_complex; in.valid = 1; in.cleartozero = 0; in.calls = calls;
	extbuffer_init_double(&o);
	o.dbuffer = out; o.buffer_max_size = (uint32_t) length; o.size_valid_elements = (uint32_t) length;
	o.valid = 1; o.cleartozero = 0; o.calls = calls;
	accummulate(&o, &in, startid, length);
}

typedef struct {
	tsdr_lib_t *tsdr;
	frameratedetector_t fd;
	extbuffer_t big, small1, small2;
	double *fp; int fcap, foff, flen;
	double *lp; int lcap, loff, llen;
	int got;
} refh_frd_t;

static void refh_frd_plot(int plot_id, int offset, double *values, int size, uint32_t samplerate, void *ctx) {
	refh_frd_t *f = (refh_frd_t *) ctx;
	(void) samplerate;
	if (plot_id == PLOT_ID_FRAME) {
		f->foff = offset; f->flen = size;
		memcpy(f->fp, values, sizeof(double) * (size < f->fcap ? size : f->fcap));
		f->got |= 1;
	} else if (plot_id == PLOT_ID_LINE) {
		f->loff = offset; f->llen = size;
		memcpy(f->lp, values, sizeof(double) * (size < f->lcap ? size : f->lcap));
		f->got |= 2;
	}
}

void *refh_frd_new(void) {
	refh_frd_t *f = (refh_frd_t *) calloc(1, sizeof(refh_frd_t));
	tsdr_init(&f->tsdr, refh_nullvalue, refh_frd_plot, f);
	f->tsdr->errormsg = NULL;
	frameratedetector_init(&f->fd, f->tsdr);
	extbuffer_init(&f->big);
	extbuffer_init_double(&f->small1);
	extbuffer_init_double(&f->small2);
	return f;
}
void refh_frd_free(void *h) {
	refh_frd_t *f = (refh_frd_t *) h;
	extbuffer_free(&f->big); extbuffer_free(&f->small1); extbuffer_free(&f->small2);
	frameratedetector_free(&f->fd);
	tsdr_free(&f->tsdr);
	free(f);
}
int refh_frd_run(void *h, uint32_t samplerate, const float *data, int size,
                 double *frame_plot, int frame_cap, int *frame_off, int *frame_len,
                 double *line_plot, int line_cap, int *line_off, int *line_len, uint64_t *calls) {
	refh_frd_t *f = (refh_frd_t *) h;
	f->fd.samplerate = samplerate;
	f->fp = frame_plot; f->fcap = frame_cap; f->lp = line_plot; f->lcap = line_cap; f->got = 0;
	frameratedetector_runontodata(&f->fd, (float *) data, size, &f->big, &f->small1, &f->small2);
	*frame_off = f->foff; *frame_len = f->flen; *line_off = f->loff; *line_len = f->llen;
	*calls = f->big.calls;
	return f->got == 3 ? 0 : 1;
}

/* ------------------------------------------------------------------ a16 (restated on the port side only) */
void refh_pixels_argb(const float *frame, int n, int inverted, const int32_t *prev, int32_t *argb) {
	(void) frame; (void) n; (void) inverted; (void) prev; (void) argb;   /* JNI glue is not buildable here */
}

/* ------------------------------------------------------------------ a22 */
void refh_complex_to_abs_diff(float *data, int size) { complex_to_abs_diff(data, size); }

int refh_superb_bestfit(const float *data1, const float *data2, int size, int samples_in_frame) {
	superbandwidth_t bw;
	superb_init(&bw);
	bw.samples_in_frame = samples_in_frame;
	const int r = superb_bestfit(&bw, (float *) data1, (float *) data2, size);
	extbuffer_free(&bw.extb_out); extbuffer_free(&bw.extb_temp);
	return r;
}

int refh_superb_ondataready(float **hops, int nhops, int count_pairs, int samples_in_frame, float *out,
                            int *best_offsets) {
	tsdr_lib_t *t;
	tsdr_init(&t, refh_nullvalue, NULL, NULL);
	t->errormsg = NULL; t->height = 1; t->refreshrate = 1.0; t->width = 0; t->pixelrate = 0;
	t->samplerate = 0; t->pixeltimeoversampletime = 0;
	superbandwidth_t bw;
	superb_init(&bw);
	bw.buffs = hops; bw.buffscount = nhops; bw.buffsbuffcount = count_pairs;
	bw.samples_in_frame = samples_in_frame; bw.samplerate = 1000; bw.alive = 1; bw.tsdr = t;
	int i;
	/* the alignment lags, obtained from the reference's own static function before the data is rotated */
	{
		const int n2 = (int) fft_getrealsize(count_pairs);
		best_offsets[0] = 0;
		for (i = 1; i < nhops; i++) best_offsets[i] = superb_bestfit(&bw, hops[0], hops[i], n2 * 2);
	}
	float *o = NULL; int osize = 0;
	superb_ondataready(&bw, &o, &osize, t);
	if (o) memcpy(out, o, sizeof(float) * 2 * (size_t) osize);
	bw.buffs = NULL;      /* caller owns the hop buffers */
	extbuffer_free(&bw.extb); extbuffer_free(&bw.extb_out); extbuffer_free(&bw.extb_temp);
	tsdr_free(&t);
	return osize;
}
