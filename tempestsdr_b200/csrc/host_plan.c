/* host_plan.c -- host-side scalar planning for libtsdrgpu.  Plain C on purpose ("host code stays C"): these are
 * the data-independent scalar recurrences of the reference, replayed with the reference's operand types and
 * operation order so the device kernels receive bit-identical parameters. */
#include "host_plan.h"
#include <math.h>

uint64_t tsdrgpu_plan_resample(double *offset, const uint32_t *sizes, uint32_t uniform, uint32_t nblocks,
                               double upsample_by, double downsample_by, tsdrgpu_rs_block_t *blocks) {
	const double r = upsample_by / downsample_by;          /* dsp.c:258 */
	const double rinv = downsample_by / upsample_by;       /* dsp.c:259 */
	double off = *offset;
	uint64_t in_pos = 0, out_pos = 0;
	for (uint32_t b = 0; b < nblocks; b++) {
		const uint32_t size = sizes ? sizes[b] : uniform;
		const uint32_t n_out = (uint32_t) (int) ((size - off) * r);      /* dsp.c:262 */
		if (n_out == 0) return UINT64_MAX;
		if (blocks) {
			blocks[b].in_start = in_pos; blocks[b].out_start = out_pos;
			blocks[b].size = size; blocks[b].n_out = n_out;
			blocks[b].r = r; blocks[b].phase = -off * r;                  /* dsp.c:272 */
		}
		off += n_out * rinv - size;                                       /* dsp.c:306 */
		in_pos += size; out_pos += n_out;
	}
	*offset = off;
	return out_pos;
}

void tsdrgpu_geometry(uint32_t samplerate, int height, double refreshrate, int *width, double *pixelrate,
                      double *pixeltimeoversampletime) {
	const double real_width = samplerate / (refreshrate * height);
	const int w = (int) (2 * real_width);
	const double pr = w * height * refreshrate;
	*width = w; *pixelrate = pr;
	*pixeltimeoversampletime = (samplerate != 0 && pr != 0) ? ((double) samplerate) / pr : 0.0;
}

/* frameratepll's write-back (syncdetector.c:141-152): `refreshrate -= vx * 1e-5` while unlocked, `avg_speed * 1e-6` once
 * locked, only when the frame moved (vx != 0).  The averages themselves come from the frame stage (fs_sync). */
int tsdrgpu_pll_step(double *refreshrate, int32_t x_vx, int32_t pll_state, double avg_speed) {
	if (x_vx == 0) return 0;
	const double diff = (pll_state == 0) ? x_vx * 0.00001 : avg_speed * 0.000001;
	*refreshrate -= diff;
	return 1;
}

void tsdrgpu_gauss_taps(float taps[5]) {
	/* CALC_GAUSSCOEFF(N,i) = expf(-2.0f*ALPHA*ALPHA*i*i/(N*N)) with textual i in {-2,-1,0,1,2}, N = 5 */
	const float e2 = expf(-2.0f * 1.0f * 1.0f * -2 * -2 / (5 * 5));
	const float e1 = expf(-2.0f * 1.0f * 1.0f * -1 * -1 / (5 * 5));
	const float e0 = expf(-2.0f * 1.0f * 1.0f * 0 * 0 / (5 * 5));
	const float norm = e2 + e1 + e0 + e1 + e2;
	taps[0] = e2 / norm; taps[1] = e1 / norm; taps[2] = e0 / norm; taps[3] = e1 / norm; taps[4] = e2 / norm;
}

/* Stage-angle errors of the reference's FFT.  fft_perform (fft.c:132-165) derives the stage twiddle c_l = (c1, c2)
 * by the half-angle recurrence c2 = sqrt((1-c1)/2), c1 = sqrt((1+c1)/2), which cancels catastrophically for small
 * angles: the angle of c_l is (pi / 2^l) * (1 + eps[l]) with |eps| growing to ~1e-6 at l = 20 and ~1e-4 at l = 23.
 * Replaying the recurrence with the same double arithmetic gives eps exactly; the CUDA FFT can then use the same
 * (wrong) angles so that its results track the reference's instead of the true DFT.  eps[0] = eps[1] = 0. */
void tsdrgpu_fft_reference_eps(int stages, int inverse, double *eps) {
	double c1 = -1.0, c2 = 0.0;
	for (int l = 0; l < stages; l++) {
		if (l == 0) eps[l] = 0.0;                      /* c = (-1, 0): only k = 0 is ever used */
		else {
			const long double ideal = 3.14159265358979323846264338327950288L / (long double) (1ull << l);
			const long double got = atan2l(fabsl((long double) c2), (long double) c1);
			eps[l] = (double) (got / ideal - 1.0L);
		}
		double n2 = sqrt((1.0 - c1) / 2.0);
		if (!inverse) n2 = -n2;
		const double n1 = sqrt((1.0 + c1) / 2.0);
		c1 = n1; c2 = n2;
	}
}

/* Auto video-mode detection, the arithmetic the GUI does on the two autocorrelation plots:
 * PlotVisualizer.java:203-236 takes the index of the first strict maximum of a plot (starting from element 0),
 * Main.java:1301-1303 turns the frame plot's into fps = samplerate / (offset + index), Main.java:1346-1350 + 1041-1043 the
 * line plot's into height = round((frame offset + index) / (line offset + index)), Java's Math.round = floor(x + 0.5). */
static int first_max(const double *v, int n) {
	int best = 0;
	double top = v[0];
	for (int i = 1; i < n; i++) if (v[i] > top) { top = v[i]; best = i; }
	return best;
}
/* the same arithmetic from peak indices that were picked elsewhere (on the device: k_plot_peaks) */
int tsdrgpu_videomode_from_peaks(int frame_offset, int frame_index, int line_offset, int line_index, uint32_t samplerate, double *fps, int *height) {
	const double frame_length = (double) (frame_offset + frame_index), line_length = (double) (line_offset + line_index);
	if (frame_length <= 0 || line_length <= 0) return TSDRGPU_EINVAL;
	if (fps) *fps = (double) (long long) samplerate / frame_length;
	if (height) *height = (int) floor(frame_length / line_length + 0.5);
	return TSDRGPU_OK;
}
int tsdrgpu_detect_videomode(const double *frame_plot, int frame_offset, int frame_len, const double *line_plot, int line_offset, int line_len,
                             uint32_t samplerate, double *fps, int *height, int *frame_index, int *line_index) {
	if (!frame_plot || !line_plot || frame_len <= 0 || line_len <= 0) return TSDRGPU_EINVAL;
	const int fi = first_max(frame_plot, frame_len), li = first_max(line_plot, line_len);
	const double frame_length = (double) (frame_offset + fi), line_length = (double) (line_offset + li);
	if (frame_length <= 0 || line_length <= 0) return TSDRGPU_EINVAL;
	if (fps) *fps = (double) (long long) samplerate / frame_length;
	if (height) *height = (int) floor(frame_length / line_length + 0.5);
	if (frame_index) *frame_index = fi;
	if (line_index) *line_index = li;
	return TSDRGPU_OK;
}
