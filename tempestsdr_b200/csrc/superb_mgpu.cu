// superb_mgpu.cu -- a22 with one hop per GPU: superb_ondataready (superbandwidth.c:121-152) sharded over H GPUs of one node.
//
// The reference aligns H hops against hop 0 (H-1 cross-correlations), transforms each hop, concatenates the spectra and runs
// one H*N-point inverse transform.  Sharded, rank q owns hop q and nothing is done twice:
//
//   phase 1  local    X_q = FFT_N(hop_q)/N,  D_q = FFT_nd(first difference of |hop_q|)/nd          -> own window
//   phase 2  lag      P = conj(D_0) D_q, IFFT_nd, grid argmax -> lag_q, written into every rank's window together with its flag
//                     (no host round trip).  D_0 comes from the rank's OWN copy of hop 0 when the caller keeps the alignment
//                     reference on every device (d_hop0: the pipeline copies the first hop to all devices while the later
//                     hops are still being recorded; 23 us of redundant transform instead of a broadcast whose source link
//                     carries (H-1) x 8 nd bytes); without it D_0 is READ FROM RANK 0 over NVLink inside the multiply, after
//                     an extra barrier (SPEC)
//   barrier  LAG      (flags in peer memory; also: every X is in its window)
//   phase 3  mix      all-to-all instead of an all-gather: rank r owns the bins m in [r N/H, (r+1) N/H).  It PULLS X_q[m] from
//                     every rank q, forms for all residues s
//                         V_s[m] = e^{2 pi i m s/(H N)} sum_q e^{2 pi i (q s/H + m lag_q/N)} X_q[m]
//                     (rotating hop q by lag_q samples == a phase ramp on its spectrum) and PUSHES V_s[m] into rank s's window:
//                     every rank receives 2 (H-1)/H N complex values instead of (H-1) N -- at H = 8, 3.5x less NVLink traffic
//   barrier  MIX
//   phase 4  residue  y[H p + s] = IDFT_N{V_s}[p] on rank s (the H N-point inverse decomposes exactly: DESIGN.md section 6),
//                     |y| (am_demod, TSDRLibrary.c:244-262: what process() does to superb_run's output) stored as float32 BY
//                     THE LAST BUTTERFLIES OF THE TRANSFORM into slot s of the ROOT rank's window (peer stores over NVLink,
//                     4 B per sample: SURVEY 8e option B) -- the transfer overlaps the transform
//   barrier  RES      (root only waits)
//   phase 5  root     interleave the H residue slots into the time-contiguous magnitude stream -> decimator -> frames
//
// Synchronisation between ranks never touches the host and never calls a collective library: every window starts with a
// small header of epoch-valued flags; a one-CTA kernel between phases stores this rank's flag into every peer (release,
// system scope, after the kernels before it have completed in stream order) and spins (acquire, bounded by a timeout that
// raises a status word instead of hanging the GPU) until every peer's flag for the phase carries this stitch's epoch.  The
// same kernels run whether the peers are other processes (windows mapped with CUDA IPC: torchrun, one process per GPU) or
// other devices of this process (cudaDeviceEnablePeerAccess: the C host library with TSDR_CUDA_DEVICES).
//
// Parity: lags exact (same float operations as the one-GPU path up to the transform, first-maximum argmax); samples are
// tolerance-based like every FFT result here (the summation order differs from the reference's radix-2 code).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int SBM_MAX_RANKS = 16;
enum { PH_SPEC = 0, PH_LAG = 1, PH_MIX = 2, PH_RES = 3, PH_COUNT = 4 };

struct WinHeader {                                   // first 4 KB of every window
	unsigned flag[PH_COUNT][SBM_MAX_RANKS];          // flag[phase][src] = epoch of the last stitch in which src finished `phase`
	int lag[SBM_MAX_RANKS];                          // lag[q] in complex samples, stored by rank q before its PH_LAG flag
	unsigned status;                                 // != 0: a wait timed out (phase + 1 in the low byte, the missing rank above it)
	unsigned pad[1024 - PH_COUNT * SBM_MAX_RANKS - SBM_MAX_RANKS - 1];
};
static_assert(sizeof(WinHeader) == 4096, "window header layout");

struct Peers { unsigned char *win[SBM_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float2 ld_peer_f2(const float2 *p) { float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ float4 ld_peer_f4(const float4 *p) { float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }

// One CTA, one thread per rank.  signal: tell every rank in `to_mask` that this rank has finished `phase` of stitch `epoch`
// (everything the stream ran before this kernel is complete; the fence + release store publish it system-wide).  wait: spin
// until every rank in `from_mask` has said the same to us.  `lag_value` (phase LAG) rides along with the flag.
__global__ void __launch_bounds__(32) sbm_sync(Peers peers, int nranks, int rank, int phase, unsigned epoch, unsigned to_mask, unsigned from_mask,
                                               const int *lag_value, long long timeout_cycles) {
	const int q = threadIdx.x;
	if (q >= nranks) return;
	WinHeader *mine = reinterpret_cast<WinHeader *>(peers.win[rank]);
	if (to_mask & (1u << q)) {
		WinHeader *theirs = reinterpret_cast<WinHeader *>(peers.win[q]);
		if (lag_value) *reinterpret_cast<volatile int *>(&theirs->lag[rank]) = *lag_value;
		__threadfence_system();
		st_release_sys(&theirs->flag[phase][rank], epoch);
	}
	if (from_mask & (1u << q)) {
		const long long t0 = clock64();
		while (ld_acquire_sys(&mine->flag[phase][q]) != epoch) {
			if (clock64() - t0 > timeout_cycles) { atomicCAS(&mine->status, 0u, (unsigned) (phase + 1) | ((unsigned) q << 8)); break; }
			__nanosleep(200);
		}
	}
}

// first difference of magnitudes (superbandwidth.c:67-81), out of place
__global__ void __launch_bounds__(256) sbm_abs_diff(const float2 *__restrict__ src, float2 *__restrict__ dst, unsigned n) {
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float2 v = src[i];
		const float cur = mag_exact(v.x, v.y);
		float prev;
		if (i == 0) prev = __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));    // seed without the square root (superbandwidth.c:70)
		else { const float2 u = src[i - 1]; prev = mag_exact(u.x, u.y); }
		dst[i] = make_float2(__fsub_rn(cur, prev), 0.0f);
	}
}

// P = (aI bI + aQ bQ, aI bQ - aQ bI) (fft.c:80-89) with a = D_0 in RANK 0's window (NVLink loads), b = D_q local
__global__ void __launch_bounds__(256) sbm_xcorr_pull(const float4 *__restrict__ d0_remote, const float4 *__restrict__ dq, float4 *__restrict__ out, unsigned n2 /* pairs of complex */) {
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
		const float4 a = ld_peer_f4(d0_remote + i), b = __ldg(dq + i);
		float4 r;
		r.x = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); r.y = __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
		r.z = __fadd_rn(__fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w)); r.w = __fsub_rn(__fmul_rn(a.z, b.w), __fmul_rn(a.w, b.z));
		out[i] = r;
	}
}

// the all-to-all mix (phase 3, see the header).  One thread per bin m of this rank's range.
struct MixArgs { Peers peers; size_t off_x, off_v; int H, log2H, rank; unsigned n; };
template <int H>
__global__ void __launch_bounds__(256) sbm_mix(MixArgs A) {
	__shared__ float2 root[H];                        // e^{2 pi i j / H}
	__shared__ int lag[H];
	if (threadIdx.x < H) {
		double sn, cs;
		sincospi(2.0 * (double) threadIdx.x / (double) H, &sn, &cs);
		root[threadIdx.x] = make_float2((float) cs, (float) sn);
		lag[threadIdx.x] = reinterpret_cast<const WinHeader *>(A.peers.win[A.rank])->lag[threadIdx.x];
	}
	__syncthreads();
	const unsigned n = A.n, per = n / H, m0 = (unsigned) A.rank * per;
	const float inv_n = 1.0f / (float) n;             // n is a power of two: exact
	for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < per; j += gridDim.x * blockDim.x) {
		const unsigned m = m0 + j;
		float2 x[H];
		#pragma unroll
		for (int q = 0; q < H; q++) x[q] = ld_peer_f2(reinterpret_cast<const float2 *>(A.peers.win[q] + A.off_x) + m);     // H loads in flight
		// z_q = e^{2 pi i m lag_q / N} X_q[m]: the argument is reduced exactly in integers before it becomes a float
		#pragma unroll
		for (int q = 0; q < H; q++) {
			const unsigned e = (unsigned) (((unsigned long long) m * (unsigned long long) (unsigned) lag[q]) & (unsigned long long) (n - 1));
			float sn, cs;
			sincospif(2.0f * ((float) e * inv_n), &sn, &cs);
			x[q] = make_float2(x[q].x * cs - x[q].y * sn, x[q].x * sn + x[q].y * cs);
		}
		// ramp e^{2 pi i m s/(H N)} = b^s, b = e^{2 pi i m/(H N)}: one double sincospi, the powers by double multiplication
		double bs, bc;
		sincospi(2.0 * ((double) m / ((double) H * (double) n)), &bs, &bc);
		double pr = 1.0, pi_ = 0.0;
		#pragma unroll
		for (int s = 0; s < H; s++) {
			float ar = 0.0f, ai = 0.0f;                // H-point DFT across the hops, sign +
			#pragma unroll
			for (int q = 0; q < H; q++) {
				const float2 w = root[(q * s) & (H - 1)];
				ar += w.x * x[q].x - w.y * x[q].y; ai += w.x * x[q].y + w.y * x[q].x;
			}
			const float rr = (float) pr, ri = (float) pi_;
			reinterpret_cast<float2 *>(A.peers.win[s] + A.off_v)[m] = make_float2(ar * rr - ai * ri, ar * ri + ai * rr);    // peer store
			const double nr = pr * bc - pi_ * bs, ni = pr * bs + pi_ * bc;
			pr = nr; pi_ = ni;
		}
	}
}

// root: stream[H p + s] = slot_s[p]
template <int H>
__global__ void __launch_bounds__(256) sbm_interleave(const float *__restrict__ slots, size_t slot_stride, unsigned n, float *__restrict__ stream) {
	for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
		float v[H];
		#pragma unroll
		for (int s = 0; s < H; s++) v[s] = __ldg(slots + (size_t) s * slot_stride + p);
		float *o = stream + (size_t) p * H;
		if (H >= 4) {
			#pragma unroll
			for (int s = 0; s < H; s += 4) *reinterpret_cast<float4 *>(o + s) = make_float4(v[s], v[s + 1], v[s + 2], v[s + 3]);
		} else *reinterpret_cast<float2 *>(o) = make_float2(v[0], v[1]);
	}
}

inline unsigned grid_for(unsigned long long n, int sm_count, int per_sm = 8) {
	const unsigned long long want = (n + 255) / 256, cap = (unsigned long long) sm_count * per_sm;
	return (unsigned) (want < cap ? (want ? want : 1) : cap);
}

}  // namespace

struct tsdrgpu_superb_mgpu {
	tsdrgpu_ctx_t *ctx;
	int H, log2H, rank, root;
	unsigned n_max;                                   // largest transform length the window was sized for
	unsigned char *win; size_t win_bytes;
	size_t off_d, off_x, off_v, off_r, slot_stride;   // byte offsets of D, X, V, residue slots; floats between two slots
	Peers peers; int connected; int ipc_opened[SBM_MAX_RANKS];
	float2 *d_work, *d_p; void *d_part; int *d_lag;   // local temporaries
	unsigned epoch; unsigned last_n;
	unsigned prep_n, prep_nd;                         // transform sizes whose kernels and twiddle tables are known to be resident
	long long timeout_cycles;
};

extern "C" {

int tsdrgpu_superb_mgpu_create(tsdrgpu_ctx_t *ctx, int nranks, int rank, int root, uint32_t max_pairs_per_hop, tsdrgpu_superb_mgpu_t **out) {
	BIND(ctx); ARG_TRY(ctx, out != NULL);
	ARG_TRY(ctx, (nranks == 2 || nranks == 4 || nranks == 8 || nranks == 16) && rank >= 0 && rank < nranks && root >= 0 && root < nranks && max_pairs_per_hop >= 64);
	tsdrgpu_superb_mgpu *g = new tsdrgpu_superb_mgpu();
	memset(g, 0, sizeof *g);
	g->ctx = ctx; g->H = nranks; g->rank = rank; g->root = root;
	while ((1 << g->log2H) < nranks) g->log2H++;
	g->n_max = tsdrgpu_fft_getrealsize(max_pairs_per_hop);
	const size_t n = g->n_max;
	g->off_d = sizeof(WinHeader);
	g->off_x = g->off_d + sizeof(float2) * n;         // nd <= n
	g->off_v = g->off_x + sizeof(float2) * n;
	g->off_r = g->off_v + sizeof(float2) * n;
	g->slot_stride = n;
	g->win_bytes = g->off_r + sizeof(float) * n * (size_t) nranks;
	CU_TRY(ctx, cudaMalloc(&g->win, g->win_bytes));
	CU_TRY(ctx, cudaMemset(g->win, 0, sizeof(WinHeader)));
	CU_TRY(ctx, cudaMalloc(&g->d_work, sizeof(float2) * n));
	CU_TRY(ctx, cudaMalloc(&g->d_p, sizeof(float2) * n));
	CU_TRY(ctx, cudaMalloc(&g->d_part, 8 * TSDRGPU_ARGMAX_PARTS));
	CU_TRY(ctx, cudaMalloc(&g->d_lag, 256));
	CU_TRY(ctx, cudaMemset(g->d_lag, 0, 256));
	CU_TRY(ctx, cudaDeviceSynchronize());
	g->peers.win[rank] = g->win;
	const char *to = getenv("TSDRGPU_SBM_TIMEOUT_MS");
	g->timeout_cycles = (long long) ((to ? atof(to) : 4000.0) * 1.9e6);      // ~1.9 GHz SM clock
	*out = g;
	return TSDRGPU_OK;
}

// unmap the peers' windows (one process per GPU: every rank disconnects, THEN every rank destroys -- a barrier of the caller's
// choosing in between -- so that no window is freed while another process still has it mapped)
int tsdrgpu_superb_mgpu_disconnect(tsdrgpu_superb_mgpu_t *g) {
	ARG_TRY((tsdrgpu_ctx_t *) NULL, g != NULL);
	BIND(g->ctx);
	CU_TRY(g->ctx, cudaDeviceSynchronize());
	for (int q = 0; q < g->H; q++) {
		if (g->ipc_opened[q] && g->peers.win[q]) { cudaIpcCloseMemHandle(g->peers.win[q]); g->ipc_opened[q] = 0; }
		if (q != g->rank) g->peers.win[q] = NULL;
	}
	g->connected = 0;
	return TSDRGPU_OK;
}

void tsdrgpu_superb_mgpu_destroy(tsdrgpu_superb_mgpu_t *g) {
	if (!g) return;
	cudaSetDevice(g->ctx->device);
	cudaDeviceSynchronize();
	for (int q = 0; q < g->H; q++) if (g->ipc_opened[q] && g->peers.win[q]) cudaIpcCloseMemHandle(g->peers.win[q]);
	cudaFree(g->win); cudaFree(g->d_work); cudaFree(g->d_p); cudaFree(g->d_part); cudaFree(g->d_lag);
	delete g;
}

int tsdrgpu_superb_mgpu_export(tsdrgpu_superb_mgpu_t *g, uint8_t handle[64]) {
	ARG_TRY((tsdrgpu_ctx_t *) NULL, g != NULL && handle != NULL);
	BIND(g->ctx);
	cudaIpcMemHandle_t h;
	CU_TRY(g->ctx, cudaIpcGetMemHandle(&h, g->win));
	memcpy(handle, &h, 64);
	return TSDRGPU_OK;
}

// one process per GPU: handles = nranks x 64 bytes, rank-major (entry `rank` is ignored)
int tsdrgpu_superb_mgpu_connect_ipc(tsdrgpu_superb_mgpu_t *g, const uint8_t *handles) {
	ARG_TRY((tsdrgpu_ctx_t *) NULL, g != NULL && handles != NULL);
	tsdrgpu_ctx_t *ctx = g->ctx;
	BIND(ctx);
	for (int q = 0; q < g->H; q++) {
		if (q == g->rank) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, handles + 64 * q, 64);
		void *p = NULL;
		CU_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		g->peers.win[q] = (unsigned char *) p; g->ipc_opened[q] = 1;
	}
	g->connected = 1;
	return TSDRGPU_OK;
}

// all ranks in this process (one device each): peer access both ways, windows cross-linked
int tsdrgpu_superb_mgpu_connect_local(tsdrgpu_superb_mgpu_t *const *all, int nranks) {
	ARG_TRY((tsdrgpu_ctx_t *) NULL, all != NULL && nranks >= 2 && nranks <= SBM_MAX_RANKS);
	for (int r = 0; r < nranks; r++) ARG_TRY((tsdrgpu_ctx_t *) NULL, all[r] != NULL && all[r]->H == nranks && all[r]->rank == r);
	for (int r = 0; r < nranks; r++) {
		tsdrgpu_ctx_t *ctx = all[r]->ctx;
		BIND(ctx);
		for (int q = 0; q < nranks; q++) {
			if (q == r) continue;
			if (all[q]->ctx->device != ctx->device) {
				int can = 0;
				CU_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, all[q]->ctx->device));
				if (!can) return tsdrgpu_fail(ctx, TSDRGPU_ENODEVICE, "no peer access between the devices of the superbandwidth group", cudaSuccess, __FILE__, __LINE__);
				const cudaError_t e = cudaDeviceEnablePeerAccess(all[q]->ctx->device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return tsdrgpu_fail(ctx, TSDRGPU_ECUDA, "cudaDeviceEnablePeerAccess", e, __FILE__, __LINE__);
				cudaGetLastError();
			}
			all[r]->peers.win[q] = all[q]->win;
		}
		all[r]->connected = 1;
	}
	return TSDRGPU_OK;
}

// Everything a stitch launches must already be loaded when the first flag-waiting kernel goes into the stream: with CUDA's lazy
// module loading the FIRST launch of a kernel may synchronise the whole context, and a context that holds a kernel spinning
// on a flag which only work not yet enqueued can raise would never drain (all ranks of one process on one device, or the
// host thread that drives every device in turn).  So the transforms of this size run once on dummy data -- which also builds
// their twiddle tables (cudaMalloc + cudaMemcpy) and sizes the scratch -- and the small kernels are touched by name.
}  // extern "C"
template <typename K> static void preload(K kernel) { cudaFuncAttributes a; cudaFuncGetAttributes(&a, kernel); cudaGetLastError(); }
extern "C" {
static int sbm_prepare(tsdrgpu_superb_mgpu *g, cudaStream_t stream, unsigned N, unsigned nd) {
	if (g->prep_n == N && g->prep_nd == nd) return TSDRGPU_OK;
	tsdrgpu_ctx_t *ctx = g->ctx;
	float2 *V = reinterpret_cast<float2 *>(g->win + g->off_v);
	int rc;
	CU_TRY(ctx, cudaMemsetAsync(g->d_work, 0, sizeof(float2) * N, stream));
	if ((rc = tsdrgpu_fft_oop_internal(ctx, stream, g->d_work, V, N, 0))) return rc;
	if ((rc = tsdrgpu_fft_oop_internal(ctx, stream, g->d_work, g->d_p, nd, 0))) return rc;
	if ((rc = tsdrgpu_fft_internal(ctx, stream, g->d_p, nd, 1))) return rc;
	if ((rc = tsdrgpu_ifft_abs_internal(ctx, stream, V, reinterpret_cast<float *>(g->d_work), N))) return rc;
	int *tmp_lag = g->d_lag + 8;
	if ((rc = tsdrgpu_argmax_mag_internal(ctx, stream, g->d_p, nd, g->d_part, tmp_lag))) return rc;
	preload(sbm_sync); preload(sbm_abs_diff); preload(sbm_xcorr_pull);
	switch (g->H) {
	case 2: preload(sbm_mix<2>); preload(sbm_interleave<2>); break;
	case 4: preload(sbm_mix<4>); preload(sbm_interleave<4>); break;
	case 8: preload(sbm_mix<8>); preload(sbm_interleave<8>); break;
	default: preload(sbm_mix<16>); preload(sbm_interleave<16>); break;
	}
	CU_TRY(ctx, cudaStreamSynchronize(stream));
	g->prep_n = N; g->prep_nd = nd;
	return TSDRGPU_OK;
}

// Rank-local part of one stitch; every rank of the group calls it once per stitch with the same count_pairs / samples_in_frame,
// each on a stream of its own device.  Asynchronous.  The root's d_stream_out receives nranks * N magnitudes (N returned in
// *h_n), time-contiguous: sample H p + s is residue s, element p.  d_stream_out is ignored on the other ranks.
int tsdrgpu_superb_mgpu_stitch(tsdrgpu_superb_mgpu_t *g, void *stream_, const float *d_hop, const float *d_hop0, int count_pairs, int samples_in_frame,
                               float *d_stream_out, uint32_t *h_n) {
	ARG_TRY((tsdrgpu_ctx_t *) NULL, g != NULL);
	tsdrgpu_ctx_t *ctx = g->ctx;
	BIND(ctx);
	ARG_TRY(ctx, g->connected && d_hop != NULL && count_pairs > 0 && samples_in_frame > 0);
	ARG_TRY(ctx, g->rank != g->root || d_stream_out != NULL);
	cudaStream_t stream = (cudaStream_t) stream_;
	const unsigned N = tsdrgpu_fft_getrealsize((uint32_t) count_pairs);
	int size = (int) ((2ull * N / (unsigned) samples_in_frame) * (unsigned) samples_in_frame);     // superbandwidth.c:84-86 with bufsize = 2N floats
	ARG_TRY(ctx, size >= 2);
	size = (int) tsdrgpu_fft_getrealsize((uint32_t) size);
	const unsigned nd = (unsigned) size / 2;
	ARG_TRY(ctx, N <= g->n_max && N >= 64 && nd >= 8 && (N % (unsigned) g->H) == 0);
	const int H = g->H, rank = g->rank;
	const unsigned all = (H >= 32) ? 0xffffffffu : ((1u << H) - 1u);
	int rc;
	if ((rc = sbm_prepare(g, stream, N, nd))) return rc;
	const unsigned epoch = ++g->epoch;
	g->last_n = N;
	float2 *D = reinterpret_cast<float2 *>(g->win + g->off_d), *X = reinterpret_cast<float2 *>(g->win + g->off_x), *V = reinterpret_cast<float2 *>(g->win + g->off_v);
	// ---- phase 1: local spectra into the own window
	if ((rc = tsdrgpu_fft_oop_internal(ctx, stream, reinterpret_cast<const float2 *>(d_hop), X, N, 0))) return rc;
	KL(ctx, "sbm_abs_diff", stream, sbm_abs_diff<<<grid_for(nd, ctx->sm_count), 256, 0, stream>>>(reinterpret_cast<const float2 *>(d_hop), g->d_work, nd));
	if ((rc = tsdrgpu_fft_oop_internal(ctx, stream, g->d_work, D, nd, 0))) return rc;
	// ---- phase 2: this rank's alignment lag against hop 0
	if (rank != 0) {
		const float4 *d0;
		if (d_hop0) {                                         // the alignment reference is resident here: its difference spectrum locally
			KL(ctx, "sbm_abs_diff", stream, sbm_abs_diff<<<grid_for(nd, ctx->sm_count), 256, 0, stream>>>(reinterpret_cast<const float2 *>(d_hop0), g->d_work, nd));
			if ((rc = tsdrgpu_fft_oop_internal(ctx, stream, g->d_work, g->d_p, nd, 0))) return rc;
			d0 = reinterpret_cast<const float4 *>(g->d_p);
		} else {                                              // pull it from rank 0 once every rank's spectra are in place
			KL(ctx, "sbm_sync", stream, sbm_sync<<<1, 32, 0, stream>>>(g->peers, H, rank, PH_SPEC, epoch, all, all, (const int *) NULL, g->timeout_cycles));
			d0 = reinterpret_cast<const float4 *>(g->peers.win[0] + g->off_d);
		}
		KL(ctx, "sbm_xcorr_pull", stream, sbm_xcorr_pull<<<grid_for(nd / 2, ctx->sm_count), 256, 0, stream>>>(d0, reinterpret_cast<const float4 *>(D), reinterpret_cast<float4 *>(g->d_p), nd / 2));
		if ((rc = tsdrgpu_fft_internal(ctx, stream, g->d_p, nd, 1))) return rc;
		if ((rc = tsdrgpu_argmax_mag_internal(ctx, stream, g->d_p, nd, g->d_part, g->d_lag))) return rc;
	} else if (!d_hop0) {                                     // rank 0 takes part in the SPEC barrier the others wait in; its lag stays 0
		KL(ctx, "sbm_sync", stream, sbm_sync<<<1, 32, 0, stream>>>(g->peers, H, rank, PH_SPEC, epoch, all, all, (const int *) NULL, g->timeout_cycles));
	}
	KL(ctx, "sbm_sync", stream, sbm_sync<<<1, 32, 0, stream>>>(g->peers, H, rank, PH_LAG, epoch, all, all, g->d_lag, g->timeout_cycles));
	// ---- phase 3: all-to-all mix
	{
		MixArgs A; A.peers = g->peers; A.off_x = g->off_x; A.off_v = g->off_v; A.H = H; A.log2H = g->log2H; A.rank = rank; A.n = N;
		const unsigned grid = grid_for(N / H, ctx->sm_count, 16);
		switch (H) {
		case 2: KL(ctx, "sbm_mix", stream, sbm_mix<2><<<grid, 256, 0, stream>>>(A)); break;
		case 4: KL(ctx, "sbm_mix", stream, sbm_mix<4><<<grid, 256, 0, stream>>>(A)); break;
		case 8: KL(ctx, "sbm_mix", stream, sbm_mix<8><<<grid, 256, 0, stream>>>(A)); break;
		default: KL(ctx, "sbm_mix", stream, sbm_mix<16><<<grid, 256, 0, stream>>>(A)); break;
		}
	}
	KL(ctx, "sbm_sync", stream, sbm_sync<<<1, 32, 0, stream>>>(g->peers, H, rank, PH_MIX, epoch, all, all, (const int *) NULL, g->timeout_cycles));
	// ---- phase 4: this rank's residue of the H N-point inverse; the last pass stores |y| straight into the root's slot
	float *slot = reinterpret_cast<float *>(g->peers.win[g->root] + g->off_r) + g->slot_stride * (size_t) rank;
	if ((rc = tsdrgpu_ifft_abs_internal(ctx, stream, V, slot, N))) return rc;
	KL(ctx, "sbm_sync", stream, sbm_sync<<<1, 32, 0, stream>>>(g->peers, H, rank, PH_RES, epoch, 1u << g->root, rank == g->root ? all : 0u, (const int *) NULL, g->timeout_cycles));
	// ---- phase 5 (root): the time-contiguous magnitude stream
	if (rank == g->root) {
		const float *slots = reinterpret_cast<const float *>(g->win + g->off_r);
		const unsigned grid = grid_for(N, ctx->sm_count, 16);
		switch (H) {
		case 2: KL(ctx, "sbm_interleave", stream, sbm_interleave<2><<<grid, 256, 0, stream>>>(slots, g->slot_stride, N, d_stream_out)); break;
		case 4: KL(ctx, "sbm_interleave", stream, sbm_interleave<4><<<grid, 256, 0, stream>>>(slots, g->slot_stride, N, d_stream_out)); break;
		case 8: KL(ctx, "sbm_interleave", stream, sbm_interleave<8><<<grid, 256, 0, stream>>>(slots, g->slot_stride, N, d_stream_out)); break;
		default: KL(ctx, "sbm_interleave", stream, sbm_interleave<16><<<grid, 256, 0, stream>>>(slots, g->slot_stride, N, d_stream_out)); break;
		}
	}
	if (h_n) *h_n = N;
	return TSDRGPU_OK;
}

// the lags (complex samples) every rank published in the last stitch, and the group's status word; synchronises `stream`
int tsdrgpu_superb_mgpu_lags(tsdrgpu_superb_mgpu_t *g, void *stream_, int *h_lags, uint32_t *h_status) {
	ARG_TRY((tsdrgpu_ctx_t *) NULL, g != NULL);
	tsdrgpu_ctx_t *ctx = g->ctx;
	BIND(ctx);
	cudaStream_t stream = (cudaStream_t) stream_;
	WinHeader h;
	CU_TRY(ctx, cudaMemcpyAsync(&h, g->win, sizeof h, cudaMemcpyDeviceToHost, stream));
	CU_TRY(ctx, cudaStreamSynchronize(stream));
	if (h_lags) for (int q = 0; q < g->H; q++) h_lags[q] = h.lag[q];
	if (h_status) *h_status = h.status;
	if (h.status) {
		char msg[160];
		snprintf(msg, sizeof msg, "superbandwidth group: rank %d never finished phase %u of the stitch (timeout)", (int) (h.status >> 8), (h.status & 0xffu) - 1u);
		return tsdrgpu_fail(ctx, TSDRGPU_ECUDA, msg, cudaSuccess, __FILE__, __LINE__);
	}
	return TSDRGPU_OK;
}

}  // extern "C"
