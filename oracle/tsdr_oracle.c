/*
 * tsdr_oracle.c -- CPU restatement of TempestSDR's IQ->raster DSP path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker, never the thing shipped or measured as the product: only tests/, the smoke() entry and
 * bench.py's cpu_baseline / --impl reference leg may load the library built from this file.
 *
 * It is a from-scratch restatement (own structure, own names) of the arithmetic the reference performs, kept
 * bit-compatible on purpose: every function names the reference file:line whose behaviour it follows, uses the
 * same operand types and the same order of floating-point operations, and is compiled with contraction off.
 * Parity status: PINNED -- tests/test_oracle_pinning.py compares every function below against the real
 * reference (oracle/_ref, built in place from /root/reference) on seeded inputs and against the committed
 * golden vectors in tests/golden/ (generated from the real reference by tests/golden/make_golden.py).
 */
#define ORC_PREFIX orc_
#include "tsdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* pixel marker values of the public API (TSDRLibrary.h:20-24) */
#define MARK_R 256.0f
#define MARK_G 512.0f
#define MARK_B 1024.0f
#define MARK_T 2048.0f

/* =====================================================================================================
 * a2  AM demodulation                                                   TSDRLibrary.c:244-262
 * magnitude of each interleaved I,Q pair, single precision, no fused multiply-add
 * ===================================================================================================== */
void orc_am_demod(const float *iq, int pairs, float *out) {
	for (int k = 0; k < pairs; k++) {
		const float re = iq[2 * k], im = iq[2 * k + 1];
		const float p = re * re;
		const float q = im * im;
		out[k] = sqrtf(p + q);
	}
}

/* =====================================================================================================
 * a6  area-weighted resampler                                            dsp.c:250-307
 *
 * The reference walks the samples once, keeping `pid` = next pixel to finish.  Sample k covers the pixel-axis
 * interval [lo, hi) with lo = k*r + phase, hi = lo + r.  Whenever a pixel ends inside the sample it is
 * emitted; what is left of the sample is banked in `contrib` for the pixel still open.
 * The output buffer is persistent between calls and only cleared the first time (extbuffer.c:47-82), which
 * matters when the loop emits fewer pixels than output_samples (e.g. r == 2 exactly): the last slot then
 * keeps whatever an earlier call left there.
 * ===================================================================================================== */
typedef struct {
	double contrib, offset;
	float *buf;
	uint32_t cap;       /* buffer_max_size */
	int never_used;     /* cleartozero */
	uint32_t last_emitted;   /* pixels the loop actually wrote in the last run (can be output_samples -+ 1) */
} orc_resampler_t;

void *orc_resample_new(void) {
	orc_resampler_t *r = (orc_resampler_t *) calloc(1, sizeof(*r));
	r->never_used = 1;
	return r;
}
void orc_resample_free(void *h) {
	orc_resampler_t *r = (orc_resampler_t *) h;
	free(r->buf);
	free(r);
}
void orc_resample_get(void *h, double *contrib, double *offset) {
	orc_resampler_t *r = (orc_resampler_t *) h;
	*contrib = r->contrib; *offset = r->offset;
}
void orc_resample_set(void *h, double contrib, double offset) {
	orc_resampler_t *r = (orc_resampler_t *) h;
	r->contrib = contrib; r->offset = offset;
}

uint32_t orc_resample_last_emitted(void *h) { return ((orc_resampler_t *) h)->last_emitted; }

/* growable-buffer rule of extbuffer_preparetohandle (extbuffer.c:47-82) */
static void resampler_fit(orc_resampler_t *r, uint32_t n) {
	if (r->cap < n || r->cap > (n << 1)) {
		/* +8: room for the reference's possible one-past-the-end emission */
		float *nb = (float *) realloc(r->buf, sizeof(float) * ((size_t) n + 8));
		r->buf = nb;
		r->cap = n;
	}
	if (r->never_used) {
		for (uint32_t i = 0; i < n; i++) r->buf[i] = 0.0f;
		r->never_used = 0;
	}
}

uint32_t orc_resample_run(void *h, const float *in, uint32_t size, double upsample_by, double downsample_by,
                          int nearest, float *out, uint32_t out_cap) {
	orc_resampler_t *st = (orc_resampler_t *) h;
	const double r = upsample_by / downsample_by;        /* pixels per sample  (dsp.c:258) */
	const double rinv = downsample_by / upsample_by;     /* samples per pixel  (dsp.c:259) */
	const uint32_t n_out = (uint32_t) (int) ((size - st->offset) * r);   /* dsp.c:262 */
	if (n_out == 0) return 0;                            /* the reference asserts here */
	resampler_fit(st, n_out);
	float *dst = st->buf;
	const double phase = -st->offset * r;                /* dsp.c:272 */

	if (nearest) {                                       /* dsp.c:274-276 */
		for (uint32_t p = 0; p < n_out; p++) dst[p] = in[((uint64_t) size * p) / n_out];
	} else {                                             /* dsp.c:280-303 */
		uint32_t open_px = 0;
		double bank = st->contrib;
		for (uint32_t k = 0; k < size; k++) {
			const double lo = k * r + phase;
			const double hi = lo + r;
			const double hi_m1 = lo + r - 1.0;
			const double v = in[k];
			if (open_px < lo && open_px < hi_m1) {       /* pixel opened earlier, closes in this sample */
				dst[open_px] = (float) (bank + v * (1.0 - lo + open_px));
				bank = 0;
				open_px++;
			}
			while (open_px < hi_m1) {                    /* pixels lying entirely inside this sample */
				dst[open_px] = (float) v;
				open_px++;
			}
			if (open_px < hi && open_px > lo) bank += (hi - open_px) * v;
			else bank += r * v;
		}
		st->contrib = bank;
		st->last_emitted = open_px;
	}
	if (nearest) st->last_emitted = n_out;
	st->offset += n_out * rinv - size;                   /* dsp.c:306 */
	memcpy(out, dst, sizeof(float) * (n_out < out_cap ? n_out : out_cap));
	return n_out;
}

/* =====================================================================================================
 * a3  block-aligned dropping                                             dsp.c:313-368
 * ===================================================================================================== */
static uint64_t drop_debt(int block, int dropped) {      /* dsp.c:321-324 */
	const uint64_t whole = (uint64_t) (dropped / block);
	return ((whole + 1) * block - dropped) % block;
}
void orc_dropcomp_shift_with(orc_dropcomp_t *s, uint32_t block, int64_t syncoffset) {   /* dsp.c:354-368 */
	if (syncoffset >= 0) s->difference -= syncoffset % block;
	else s->difference -= block + syncoffset % block;
	if (s->difference < 0) s->difference = (int64_t) drop_debt((int) block, (int) -s->difference);
}
int orc_dropcomp_will_drop_all(orc_dropcomp_t *s, uint32_t size, uint32_t block) {      /* dsp.c:348-352 */
	(void) block;
	return size <= s->difference;
}
uint32_t orc_dropcomp_add(orc_dropcomp_t *s, uint32_t size, uint32_t block, int ring_accepts, uint32_t *skip) {
	/* dsp.c:326-346 */
	*skip = 0;
	if (size <= s->difference) { s->difference -= size; *skip = size; return 0; }
	if (ring_accepts) {
		const uint32_t lead = (uint32_t) s->difference;
		s->difference = 0;
		*skip = lead;
		return size - lead;
	}
	s->difference -= size % block;
	if (s->difference < 0) s->difference = (int64_t) drop_debt((int) block, (int) -s->difference);
	return 0;
}

/* =====================================================================================================
 * a4  geometry                                                           TSDRLibrary.c:540-550
 * ===================================================================================================== */
void orc_geometry(uint32_t samplerate, int height, double refreshrate, int *width, double *pixelrate,
                  double *pixeltimeoversampletime) {
	const double real_width = samplerate / (refreshrate * height);
	const int w = (int) (2 * real_width);
	const double prate = w * height * refreshrate;
	*width = w; *pixelrate = prate;
	*pixeltimeoversampletime = (samplerate != 0 && prate != 0) ? ((double) samplerate) / prate : 0.0;
}

/* =====================================================================================================
 * a8  auto-gain                                                          dsp.c:41-94
 * ===================================================================================================== */
static int px_is_marker(float v) { return v > 250.0 || v < -250; }      /* dsp.c:57 */

void orc_autogain(orc_autogain_t *s, int n, const float *in, float *out, float norm) {
	float lo = in[0], hi = in[0];
	double total = 0.0;
	for (int i = 0; i < n; i++) {
		const float v = in[i];
		if (px_is_marker(v)) continue;
		if (v > hi) hi = v; else if (v < lo) lo = v;
		total += v;
	}
	const float keep = 1.0f - norm;
	s->lastmax = keep * s->lastmax + norm * hi;
	s->lastmin = keep * s->lastmin + norm * lo;
	const float span = (s->lastmax == s->lastmin) ? 1.0f : (s->lastmax - s->lastmin);

	const double mean = total / (double) n;
	double sq = 0.0, lin = 0.0;
	for (int i = 0; i < n; i++) {
		const float v = in[i];
		out[i] = px_is_marker(v) ? v : ((v - s->lastmin) / span);
		const double d = v - mean;
		sq += d * d;
		lin += d;
	}
	const double stdev = sqrt((sq - lin * lin / (double) n) / (double) (n - 1));
	s->snr = (float) (mean / stdev);
}

/* =====================================================================================================
 * a9  temporal IIR ("motion blur")                                       dsp.c:22-33
 * first product in float, second in double, sum in double, store float
 * ===================================================================================================== */
void orc_timelowpass(float coeff, int n, const float *in, float *screen) {
	const double fresh = 1.0 - coeff;
	for (int i = 0; i < n; i++) {
		const float old = screen[i] * coeff;
		screen[i] = (float) (old + in[i] * fresh);
	}
}

/* =====================================================================================================
 * a10 column / row collapse                                              dsp.c:96-110
 * sequential single-precision accumulation in raster order
 * ===================================================================================================== */
void orc_average_v_h(int w, int h, const float *in, float *wbuf, float *hbuf) {
	for (int x = 0; x < w; x++) wbuf[x] = 0.0f;
	for (int y = 0; y < h; y++) hbuf[y] = 0.0f;
	for (int y = 0; y < h; y++) {
		const float *row = in + (size_t) y * w;
		for (int x = 0; x < w; x++) {
			wbuf[x] += row[x];
			hbuf[y] += row[x];
		}
	}
}

/* =====================================================================================================
 * a14 5-tap circular Gaussian                                            gaussian.c:18-79
 * For n >= 5 the in-place sliding-window code of the reference equals an out-of-place circular convolution
 * out[j] = sum_{k=-2..2} in[(j+k) mod n] * c[k], summed left to right.  For n < 5 the reference's window
 * bookkeeping degenerates; it is replayed literally.
 * ===================================================================================================== */
static void gauss_taps(float c[5]) {
	/* exp(-2 * i^2 / 25), i = -2..2, normalised by their single-precision sum (gaussian.c:16-30) */
	const float g2 = expf(-2.0f * 1.0f * 1.0f * -2 * -2 / (5 * 5));
	const float g1 = expf(-2.0f * 1.0f * 1.0f * -1 * -1 / (5 * 5));
	const float g0 = expf(-2.0f * 1.0f * 1.0f * 0 * 0 / (5 * 5));
	const float sum = g2 + g1 + g0 + g1 + g2;
	c[0] = g2 / sum; c[1] = g1 / sum; c[2] = g0 / sum; c[3] = g1 / sum; c[4] = g2 / sum;
}

void orc_gaussianblur(float *data, int n) {
	float c[5];
	gauss_taps(c);
	if (n >= 5) {
		float *src = (float *) malloc(sizeof(float) * n);
		memcpy(src, data, sizeof(float) * n);
		for (int j = 0; j < n; j++) {
			const int a = (j + n - 2) % n, b = (j + n - 1) % n, d = (j + 1) % n, e = (j + 2) % n;
			data[j] = src[a] * c[0] + src[b] * c[1] + src[j] * c[2] + src[d] * c[3] + src[e] * c[4];
		}
		free(src);
		return;
	}
	/* literal replay for tiny strips (gaussian.c:32-78) */
	float w0 = data[0], w1 = data[1 % n], w2 = data[2 % n], w3 = data[3 % n], w4 = data[4 % n];
	const float keep2 = w2, keep3 = w3, keep4 = w4;
	for (int i = 0; i < n; i++) {
		const int upd = (i < n - 2) ? (i + 2) : (i - (n - 2));
		const int nxt = (i < n - 5) ? (i + 5) : (i - (n - 5));
		data[upd] = w0 * c[0] + w1 * c[1] + w2 * c[2] + w3 * c[3] + w4 * c[4];
		w0 = w1; w1 = w2; w2 = w3; w3 = w4;
		if (nxt < 2 || nxt >= 5) w4 = data[nxt];
		else w4 = (nxt == 2) ? keep2 : (nxt == 3 ? keep3 : keep4);
	}
}

/* =====================================================================================================
 * a13 best blanking-strip position for one strip width                   syncdetector.c:26-58
 * score(window) = (mean outside - mean inside)^2 ; first maximum wins ; the recorded index is the loop
 * counter BEFORE the window slides (syncdetector.c:53-56)
 * ===================================================================================================== */
void orc_findbestfit(const float *data, int size, float totalsum, int stripsize, double *bestfit, int *bestfitid) {
	const double n_out = size - stripsize;
	const double n_in = stripsize;
	double inside = 0.0;
	for (int i = 0; i < stripsize; i++) inside += data[i];
	double contrast = (totalsum - inside) / n_out - inside / n_in;
	double best = contrast * contrast;
	int where = 0;
	const int wrap_at = size - stripsize;
	for (int i = 0; i < size - 1; i++) {
		const double leaving = data[i];
		const int enter_idx = (i < wrap_at) ? (i + stripsize) : (i - wrap_at);
		const double entering = data[enter_idx];
		inside = inside - leaving + entering;
		contrast = (totalsum - inside) / n_out - inside / n_in;
		const double score = contrast * contrast;
		if (score > best) { best = score; where = i; }
	}
	*bestfit = best; *bestfitid = where;
}

/* =====================================================================================================
 * a12 blanking-interval search + smoothed centre                         syncdetector.c:71-119
 * ===================================================================================================== */
void orc_findthesweetspot(orc_sweetspot_t *s, float *data, int size, int minsize, double lowpasscoeff) {
	if (minsize < 1) minsize = 1;
	const int half = size >> 1;
	if (s->curr_stripsize < minsize) s->curr_stripsize = minsize;
	else if (s->curr_stripsize > half) s->curr_stripsize = half;

	orc_gaussianblur(data, size);

	double total = 0.0;
	for (int i = 0; i < size; i++) total += data[i];
	const float totalf = (float) total;                  /* findbestfit takes a float (syncdetector.c:26) */

	const int base = s->curr_stripsize;
	int best_size = base, best_start;
	double best_score;
	orc_findbestfit(data, size, totalf, base, &best_score, &best_start);

	const int tries[4] = { base - 4, base + 4, base >> 1, base << 1 };    /* syncdetector.c:90-93 */
	for (int t = 0; t < 4; t++) {
		const int cand = tries[t];
		if (cand >= minsize && cand < half && cand != base) {
			double sc; int st;
			orc_findbestfit(data, size, totalf, cand, &sc, &st);
			if (sc > best_score) { best_score = sc; best_start = st; best_size = cand; }
		}
	}
	s->curr_stripsize = best_size;

	data[best_start] = MARK_B;                           /* syncdetector.c:98-99 */
	data[(best_start + best_size) % size] = MARK_B;

	const int h2 = size / 2;
	int centre = (best_start + best_size / 2) % size;
	const int jump = centre - s->dx;
	if (jump > h2) s->dx += size;
	else if (jump < -h2) centre += size;

	const int before = s->dx;
	s->dx = (int) (((int64_t) round(centre * lowpasscoeff + (1.0 - lowpasscoeff) * s->dx)) % ((int64_t) size));
	const int moved = s->dx - before;
	s->vx = (moved > h2) ? (size - moved) : ((moved < -h2) ? (-size - moved) : moved);
	s->absvx = (s->vx >= 0) ? s->vx : -s->vx;
}

/* =====================================================================================================
 * a7 / a11 / a15  the per-frame stage                     dsp.c:112-239, syncdetector.c:133-226
 * ===================================================================================================== */
typedef struct {
	orc_pp_config_t cfg;
	double refreshrate;      /* live copy, moved by the PLL */
	int width_live;          /* tsdr->width as set_internal_samplerate leaves it */
	/* dsp_postprocess_t */
	float *screen, *send, *corrected, *wbuf, *hbuf;
	int n, w, h, cap, runs, lp_before_sync;
	orc_autogain_t ag;
	/* syncdetector_t */
	orc_sweetspot_t sx, sy;
	double avg_speed;
	int pll_state;
} orc_pp_t;

void *orc_pp_new(void) {
	orc_pp_t *p = (orc_pp_t *) calloc(1, sizeof(*p));
	p->ag.snr = 1.0f;        /* dsp.c:35-39 */
	p->refreshrate = 1.0; p->cfg.height = 1; p->cfg.refreshrate = 1.0;
	return p;
}
void orc_pp_free(void *h) {
	orc_pp_t *p = (orc_pp_t *) h;
	free(p->screen); free(p->send); free(p->corrected); free(p->wbuf); free(p->hbuf);
	free(p);
}
void orc_pp_config(void *h, const orc_pp_config_t *cfg) {
	orc_pp_t *p = (orc_pp_t *) h;
	p->cfg = *cfg;
	p->refreshrate = cfg->refreshrate;
	double pr, ptos;
	orc_geometry(cfg->samplerate, cfg->height, cfg->refreshrate, &p->width_live, &pr, &ptos);
}

static void draw_vline(int x, float *d, int w, int h, float v) { for (int y = 0; y < h; y++) d[x + w * y] = v; }
static void draw_hline(int y, float *d, int w, int h, float v) { (void) h; for (int x = 0; x < w; x++) d[x + w * y] = v; }

/* syncdetector_run (syncdetector.c:171-226); returns the buffer holding the result */
static float *pp_sync(orc_pp_t *p, orc_pp_result_t *res, float *data, float *outdata, int w, int h,
                      int greenlines, int may_modify) {
	orc_findthesweetspot(&p->sx, p->wbuf, w, (int) (w * 0.05f), 0.9);
	orc_findthesweetspot(&p->sy, p->hbuf, h, (int) (h * 0.01f), 0.1);

	/* frameratepll (syncdetector.c:133-153) */
	p->avg_speed = p->avg_speed * 0.99 + 0.01 * p->sx.vx;
	p->pll_state = (p->avg_speed < 0.5 && p->avg_speed > -0.5) ? 1 : 0;
	if (p->cfg.pll && p->sx.vx != 0) {
		const double step = (p->pll_state == 0) ? p->sx.vx * 0.00001 : p->avg_speed * 0.000001;
		p->refreshrate -= step;
		double pr, ptos;
		orc_geometry(p->cfg.samplerate, p->cfg.height, p->refreshrate, &p->width_live, &pr, &ptos);
		if (res) res->pll_callback_fired = 1;
	}

	if (p->cfg.autoshift) {                              /* circular 2-D re-centre, syncdetector.c:187-207 */
		const int dx = p->sx.dx, dy = p->sy.dx;
		for (int y = 0; y < h; y++) {
			const int sy_ = (y + dy) % h;                /* output row y comes from input row y+dy */
			const float *srow = data + (size_t) sy_ * w;
			float *drow = outdata + (size_t) y * w;
			memcpy(drow, srow + dx, sizeof(float) * (size_t) (w - dx));
			memcpy(drow + (w - dx), srow, sizeof(float) * (size_t) dx);
		}
		return outdata;
	}
	if (greenlines && may_modify) {
		draw_vline(p->sx.dx, data, w, h, MARK_G);
		draw_hline(p->sy.dx, data, w, h, MARK_G);
		return data;
	}
	if (greenlines) {
		memcpy(outdata, data, sizeof(float) * (size_t) w * h);
		draw_vline(p->sx.dx, outdata, w, h, MARK_G);
		draw_hline(p->sy.dx, outdata, w, h, MARK_G);
		return outdata;
	}
	return data;
}

int orc_pp_run(void *hnd, const float *frame_in, int w, int hgt, float motionblur, float lowpasscoeff,
               int lowpass_before_sync, int autogain_after_proc, float *frame_out, orc_pp_result_t *res) {
	orc_pp_t *p = (orc_pp_t *) hnd;
	if (res) memset(res, 0, sizeof(*res));

	/* buffer (re)sizing, dsp.c:152-173: only the screen buffer is zeroed, and only when it grows */
	if (hgt != p->h || w != p->w) {
		const int oldw = p->w, oldh = p->h;
		p->h = hgt; p->w = w; p->n = w * hgt;
		if (p->n > p->cap) {
			p->cap = p->n;
			p->screen = (float *) realloc(p->screen, sizeof(float) * p->cap);
			p->send = (float *) realloc(p->send, sizeof(float) * p->cap);
			p->corrected = (float *) realloc(p->corrected, sizeof(float) * p->cap);
			for (int i = 0; i < p->cap; i++) p->screen[i] = 0.0f;
		}
		if (w != oldw) p->wbuf = (float *) realloc(p->wbuf, sizeof(float) * w);
		if (hgt != oldh) p->hbuf = (float *) realloc(p->hbuf, sizeof(float) * hgt);
	}
	if (p->lp_before_sync != lowpass_before_sync) {      /* dsp.c:178-186 */
		p->lp_before_sync = lowpass_before_sync;
		for (int i = 0; i < p->n; i++) { p->screen[i] = 0.0f; p->send[i] = 0.0f; p->corrected[i] = 0.0f; }
	}

	float *scratch = (float *) malloc(sizeof(float) * p->n);
	memcpy(scratch, frame_in, sizeof(float) * p->n);
	float *input = scratch;
	if (!autogain_after_proc) {
		orc_autogain(&p->ag, p->n, input, p->send, lowpasscoeff);
		input = p->send;
	}
	float *result;
	if (lowpass_before_sync) {                           /* dsp.c:201-212 */
		orc_timelowpass(motionblur, p->n, input, p->screen);
		orc_average_v_h(p->w, p->h, p->screen, p->wbuf, p->hbuf);
		float *synced = pp_sync(p, res, p->screen, p->corrected, p->w, p->h, !p->cfg.superres, 0);
		if (autogain_after_proc) { orc_autogain(&p->ag, p->n, synced, p->send, lowpasscoeff); result = p->send; }
		else result = synced;
	} else {                                             /* dsp.c:214-226 */
		orc_average_v_h(p->w, p->h, input, p->wbuf, p->hbuf);
		float *synced = pp_sync(p, res, input, p->corrected, p->w, p->h,
		                        (motionblur == 0.0f) && (!p->cfg.superres), 1);
		orc_timelowpass(motionblur, p->n, synced, p->screen);
		if (autogain_after_proc) { orc_autogain(&p->ag, p->n, p->screen, p->send, lowpasscoeff); result = p->send; }
		else result = p->screen;
	}
	memcpy(frame_out, result, sizeof(float) * p->n);
	free(scratch);

	int ag_fired = 0;
	if (p->runs++ > 5) { p->runs = 0; ag_fired = 1; }    /* dsp.c:231-235 */
	if (res) {
		res->x = p->sx; res->y = p->sy;
		res->avg_speed = p->avg_speed; res->pll_state = p->pll_state;
		res->lastmax = p->ag.lastmax; res->lastmin = p->ag.lastmin; res->snr = p->ag.snr;
		res->refreshrate_after = p->refreshrate; res->width_after = p->width_live;
		res->autogain_callback_fired = ag_fired;
		if (ag_fired) { res->autogain_cb_min = p->ag.lastmin; res->autogain_cb_max = p->ag.lastmax; }
	}
	return 0;
}

/* =====================================================================================================
 * a16 host pixel rule (what "integer pixel" means for parity)        JavaGUI/jni/TSDRLibraryNDK.c:222-283
 * ===================================================================================================== */
void orc_pixels_argb(const float *frame, int n, int inverted, const int32_t *prev, int32_t *argb) {
	const int32_t white = 255 | (255 << 8) | (255 << 16);
	for (int i = 0; i < n; i++) {
		const float v = frame[i];
		int32_t px;
		if (v > 0.0f && v <= 1.0f) {
			int g = (int) (v * 255.0f);
			if (inverted) g = 255 - g;
			px = g | (g << 8) | (g << 16);
		} else if (v <= 0.0f) px = inverted ? white : 0;
		else if (v == MARK_R) px = 255 << 16;
		else if (v == MARK_G) px = 255 << 8;
		else if (v == MARK_B) px = 255;
		else if (v == MARK_T) px = prev ? prev[i] : 0;   /* slot left untouched */
		else px = inverted ? 0 : white;
		argb[i] = px;
	}
}

/* =====================================================================================================
 * a20 radix-2 FFT: float storage, double arithmetic, twiddles by recurrence     fft.c:96-176
 * ===================================================================================================== */
uint32_t orc_fft_getrealsize(uint32_t size) {            /* fft.c:5-11: largest power of two <= size */
	uint32_t bits = 0;
	while ((size /= 2) != 0) bits++;
	return 1u << bits;
}

static void bit_reverse_permute(float *iq, uint32_t n) {
	uint32_t j = 0;
	for (uint32_t i = 0; i + 1 < n; i++) {
		if (i < j) {
			const float tr = iq[2 * i], ti = iq[2 * i + 1];
			iq[2 * i] = iq[2 * j]; iq[2 * i + 1] = iq[2 * j + 1];
			iq[2 * j] = tr; iq[2 * j + 1] = ti;
		}
		uint32_t bit = n >> 1;
		while (bit != 0 && bit <= j) { j -= bit; bit >>= 1; }
		j += bit;
	}
}

void orc_fft(float *iq, uint32_t size, int inverse) {
	int stages = 0;
	while ((size /= 2) != 0) stages++;
	const uint32_t n = 1u << stages;
	bit_reverse_permute(iq, n);

	double step_re = -1.0, step_im = 0.0;                /* e^{-+ i pi / half}, refined by half-angle */
	uint32_t span = 1;
	for (int s = 0; s < stages; s++) {
		const uint32_t half = span;
		span <<= 1;
		double tw_re = 1.0f, tw_im = 0.0f;
		for (uint32_t j = 0; j < half; j++) {
			for (uint32_t top = j; top < n; top += span) {
				const uint32_t bot = top + half;
				const double br = iq[2 * bot], bi = iq[2 * bot + 1];
				const double pr = tw_re * br - tw_im * bi;
				const double pi = tw_re * bi + tw_im * br;
				iq[2 * bot]     = (float) (iq[2 * top] - pr);
				iq[2 * bot + 1] = (float) (iq[2 * top + 1] - pi);
				iq[2 * top]     = (float) (iq[2 * top] + pr);
				iq[2 * top + 1] = (float) (iq[2 * top + 1] + pi);
			}
			const double nr = tw_re * step_re - tw_im * step_im;
			tw_im = tw_re * step_im + tw_im * step_re;
			tw_re = nr;
		}
		step_im = sqrt((1.0 - step_re) / 2.0);
		if (!inverse) step_im = -step_im;
		step_re = sqrt((1.0 + step_re) / 2.0);
	}
	if (!inverse) {
		const float scale = (float) n;
		for (uint32_t i = 0; i < 2 * n; i++) iq[i] /= scale;
	}
}

/* a19  "autocorrelation" = IFFT(|FFT(x)| / N)                            fft.c:49-64
 * all `size` samples are widened to complex and all `size` bins get abs(); only the first
 * N = 2^floor(log2 size) take part in the transforms */
void orc_autocorrelation(float *answer, const float *real, uint32_t size) {
	for (uint32_t i = 0; i < size; i++) { answer[2 * i] = real[i]; answer[2 * i + 1] = 0.0f; }
	const uint32_t n = orc_fft_getrealsize(size);
	orc_fft(answer, n, 0);
	for (uint32_t i = 0; i < size; i++) {
		const float re = answer[2 * i], im = answer[2 * i + 1];
		const float p = re * re, q = im * im;
		answer[2 * i] = sqrtf(p + q);
		answer[2 * i + 1] = 0;
	}
	orc_fft(answer, n, 1);
}

/* cross-correlation through the spectrum: IFFT( A * conj-ish(B) )       fft.c:69-93 */
void orc_crosscorrelation(float *a_out, float *b_tmp, uint32_t samples) {
	const uint32_t n = orc_fft_getrealsize(samples);
	orc_fft(a_out, n, 0);
	orc_fft(b_tmp, n, 0);
	for (uint32_t i = 0; i < n; i++) {
		const float ar = a_out[2 * i], ai = a_out[2 * i + 1];
		const float br = b_tmp[2 * i], bi = b_tmp[2 * i + 1];
		const float t0 = ar * br, t1 = ai * bi, t2 = ar * bi, t3 = ai * br;
		a_out[2 * i] = t0 + t1;
		a_out[2 * i + 1] = t2 - t3;
	}
	orc_fft(a_out, n, 1);
}

/* =====================================================================================================
 * a18 running mean of lag magnitudes                           frameratedetector.c:34-62, 87-126
 * ===================================================================================================== */
void orc_accumulate(double *out, uint64_t calls, const float *in_complex, int startid, int length) {
	const float *src = in_complex + (size_t) startid * 2;
	const double now_n = (double) calls, before_n = (double) (calls - 1);
	for (int i = 0; i < length; i++) {
		const double re = src[2 * i], im = src[2 * i + 1];
		const double mag = sqrt(re * re + im * im);
		out[i] = (calls == 0) ? mag : (out[i] * before_n + mag) / now_n;
	}
}

void orc_framerate_windows(uint32_t samplerate, int *frame_min, int *frame_max, int *line_min, int *line_max) {
	/* frameratedetector.c:91-95 with MIN/MAX_FRAMERATE 55/87, MIN/MAX_HEIGHT 590/1500 */
	*frame_max = (int) (samplerate / (double) (55));
	*frame_min = (int) (samplerate / (double) (87));
	*line_max = (int) (samplerate / (double) (590 * 55));
	*line_min = (int) (samplerate / (double) (1500 * 87));
}
uint32_t orc_framerate_capture_size(uint32_t samplerate) {   /* frameratedetector.c:160 */
	return (uint32_t) (3.1 * samplerate / (double) (55));
}

typedef struct {
	float *big; uint32_t big_cap; uint64_t calls; int fresh;
	double *p1, *p2; uint32_t p1_cap, p2_cap;
} orc_frd_t;

void *orc_frd_new(void) {
	orc_frd_t *f = (orc_frd_t *) calloc(1, sizeof(*f));
	f->fresh = 1;
	return f;
}
void orc_frd_free(void *h) {
	orc_frd_t *f = (orc_frd_t *) h;
	free(f->big); free(f->p1); free(f->p2); free(f);
}
static void plot_fit(double **buf, uint32_t *cap, uint32_t n, int fresh) {
	if (*cap < n || *cap > (n << 1)) { *buf = (double *) realloc(*buf, sizeof(double) * n); *cap = n; }
	if (fresh) for (uint32_t i = 0; i < n; i++) (*buf)[i] = 0.0;
}
int orc_frd_run(void *h, uint32_t samplerate, const float *data, int size,
                double *frame_plot, int frame_cap, int *frame_off, int *frame_len,
                double *line_plot, int line_cap, int *line_off, int *line_len, uint64_t *calls) {
	orc_frd_t *f = (orc_frd_t *) h;
	int fmin, fmax, lmin, lmax;
	orc_framerate_windows(samplerate, &fmin, &fmax, &lmin, &lmax);
	const uint32_t need = 2u * (uint32_t) size;
	if (f->big_cap < need || f->big_cap > (need << 1)) {
		f->big = (float *) realloc(f->big, sizeof(float) * need); f->big_cap = need;
	}
	if (f->fresh) f->calls = 0;
	f->calls++;                                          /* extbuffer.c:81: one prepare per capture */
	orc_autocorrelation(f->big, data, (uint32_t) size);
	plot_fit(&f->p1, &f->p1_cap, (uint32_t) (fmax - fmin), f->fresh);
	plot_fit(&f->p2, &f->p2_cap, (uint32_t) (lmax - lmin), f->fresh);
	f->fresh = 0;
	orc_accumulate(f->p1, f->calls, f->big, fmin, fmax - fmin);
	orc_accumulate(f->p2, f->calls, f->big, lmin, lmax - lmin);
	*frame_off = fmin; *frame_len = fmax - fmin; *line_off = lmin; *line_len = lmax - lmin;
	memcpy(frame_plot, f->p1, sizeof(double) * (size_t) ((fmax - fmin) < frame_cap ? (fmax - fmin) : frame_cap));
	memcpy(line_plot, f->p2, sizeof(double) * (size_t) ((lmax - lmin) < line_cap ? (lmax - lmin) : line_cap));
	*calls = f->calls;
	return 0;
}

/* =====================================================================================================
 * a22 superbandwidth stitch                                              superbandwidth.c:67-152
 * ===================================================================================================== */
void orc_complex_to_abs_diff(float *data, int size) {    /* superbandwidth.c:67-81 */
	/* note: the seed is the squared magnitude of pair 0, without the square root */
	float before = data[0] * data[0] + data[1] * data[1];
	for (int i = 0; i < size; i += 2) {
		const float re = data[i], im = data[i + 1];
		const float p = re * re, q = im * im;
		const float now = sqrtf(p + q);
		data[i] = now - before;
		data[i + 1] = 0;
		before = now;
	}
}

int orc_superb_bestfit(const float *data1, const float *data2, int size, int samples_in_frame) {
	/* superbandwidth.c:83-119 */
	size = (size / samples_in_frame) * samples_in_frame;
	size = (int) orc_fft_getrealsize((uint32_t) size);
	const int pairs = size / 2;
	float *a = (float *) malloc(sizeof(float) * size);
	float *b = (float *) malloc(sizeof(float) * size);
	memcpy(a, data1, sizeof(float) * size);
	memcpy(b, data2, sizeof(float) * size);
	orc_complex_to_abs_diff(a, size);
	orc_complex_to_abs_diff(b, size);
	orc_crosscorrelation(a, b, (uint32_t) pairs);
	int lag = 0;
	float peak = 0.0f;
	for (int i = 0; i < pairs; i++) {
		const float re = a[2 * i], im = a[2 * i + 1];
		const float p = re * re, q = im * im;
		const float m = sqrtf(p + q);
		if (i == 0) peak = m;
		else if (m > peak) { peak = m; lag = i; }
	}
	free(a); free(b);
	return 2 * lag;
}

int orc_superb_ondataready(float **hops, int nhops, int count_pairs, int samples_in_frame, float *out,
                           int *best_offsets) {
	/* superbandwidth.c:121-152 */
	const uint32_t n = orc_fft_getrealsize((uint32_t) count_pairs);
	const int floats = (int) (n * 2);
	float *tmp = (float *) malloc(sizeof(float) * floats);
	best_offsets[0] = 0;
	for (int i = 1; i < nhops; i++) {
		const int shift = orc_superb_bestfit(hops[0], hops[i], floats, samples_in_frame);
		best_offsets[i] = shift;
		/* rotate left by `shift` floats */
		memcpy(tmp, hops[i] + shift, sizeof(float) * (size_t) (floats - shift));
		memcpy(tmp + (floats - shift), hops[i], sizeof(float) * (size_t) shift);
		memcpy(hops[i], tmp, sizeof(float) * (size_t) floats);
		orc_fft(hops[i], n, 0);
	}
	orc_fft(hops[0], n, 0);
	for (int i = 0; i < nhops; i++) memcpy(out + (size_t) i * floats, hops[i], sizeof(float) * (size_t) floats);
	orc_fft(out, n * (uint32_t) nhops, 1);
	free(tmp);
	return (int) (n * (uint32_t) nhops);
}
