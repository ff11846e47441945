// tvq_fwd_stream.cuh — fused VQ forward with tcgen05 scoring and a STREAMED codebook: any k, d <= 256.
// This is the path of BASELINE configs[2] (k = 512 ... 16384, d = 64 ... 256, millions of latents),
// where the distance computation is a real GEMM (2*k*d FLOP per latent) and, from k ~ 1000 up,
// the tensor pipe — not HBM — is the roofline.
//
// Scoring runs in bf16 on the tensor cores (fp32 accumulate in tensor memory); the bf16 scores only
// NOMINATE candidates.  The decision itself is the canonical fp32/fp64 rule shared by every path
// (oracle/vq_canon.c): a code can be the canonical arg-min only if its bf16
// score lies within a rigorous error bound of the best bf16 score, so those codes (1.2 ... 1.8 per
// row on Gaussian data) are re-scored in fp32 and, if still inseparable, in fp64.  Indices are
// therefore bit-identical to the CUDA-core path and to the C oracle.
//
// One persistent CTA per SM, 12 or 20 warps, warp-specialised; a row tile is 128 latents (UMMA M = 128:
// TMEM lane = latent), a code tile is NT codes (one tcgen05.mma N), d is cut into 64-column slabs:
//   warp 0      producer: TMA (cp.async.bulk.tensor, SWIZZLE_128B) of bf16 code slabs [NT x 64] of the NEGATED codebook
//               into a ring of shared-memory stages, then, per code tile, the [NT x 16] slab of |e|^2 / 2 pieces (SWIZZLE_32B)
//   warp 1      MMA issuer: per code tile d/16 tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32, M = 128, N = NT) plus ONE more
//               K step — constant A rows (1, 1, 1, 0...) against the three bf16 pieces of |e_c|^2 / 2 — into one of 512/NT
//               tensor-memory score slots: the accumulator IS the half-score |e|^2 / 2 - x.e; tcgen05.commit.
//               (Producer and issuer are warp-uniform loops with one elected lane: uniform-register descriptors, carried
//               stage index / phase — under `if (lane == 0)` the issuer was instruction-bound.)
//   warps 2-3   converters: x rows fp32 -> bf16 A operand in UMMA K-major SWIZZLE_128B layout, double buffered; at d <= 128
//               the rows arrive through shared memory (8 KB blocks, cp.async.bulk issued two blocks ahead by the warp
//               itself); per latent |x|^2 and |x - bf16(x)|^2, from which the scan thread of the latent takes its bound.
//   warps 4-..  scan + apply, SP = 2 or 4 warps per TMEM lane quadrant (each takes 1/SP of the columns of every code tile),
//               ONE THREAD PER LATENT: the running minimum m is SEEDED from a recording-free pass over the first code tile
//               and shared by the parts of a quadrant; then tcgen05.ld of 32 scores at a time (the score slot is handed
//               back as soon as the warp's last load of the code tile has completed), 3-input-min tree, threshold m + bound;
//               a 32-score chunk is looked at again only if its minimum beats the threshold: a straight-line 8-compare
//               mask of the groups of 4 that hold a hit, then one indexed branch per hit group; the candidates go to a
//               small per-latent list in shared memory.  After the last code tile the parts merge their minima, every
//               latent gets a record (count + up to four candidates), and each warp resolves 32/SP latents: one-candidate
//               latents need no decision (x row and code word of all of them in flight at once), ambiguous ones take the
//               cascade above; then idx / q (straight-through) / loss partial / EMA statistics (red.global.v4).
// For k >= 512 the CTAs work in PAIRS (template parameter CG = 2: 2-CTA clusters, tcgen05 cta_group::2): see
// the comment at the kernel.
// Algorithmic cost per latent: 8d + 8 bytes of HBM (x is re-read once from L2 by the apply phase),
// 2*k*d tensor FLOP.
// Nothing in the hot loops may spill: with ~225 KB of shared memory the L1 is gone and local memory lives in L2.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "tvq_common.cuh"
#include "tvq_fwd_simt.cuh"
#include "tvq_sm100.cuh"

namespace tvq {

#ifdef TVQ_STREAM_PROF
// Experiment build only (tools/profile_stream.py): CTA 0's per-role wait / work clock totals.
__device__ unsigned long long g_sprof[128];
#define SP_DECL long long sp[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long sp_t = clock64(); (void)sp_t
#define SP_WAIT(i, bar, par) do { long long _t = clock64(); mbar_wait(bar, par); sp[i] += clock64() - _t; } while (0)
#define SP_WAITL(i, bar, par) do { long long _t = clock64(); mbar_wait_long(bar, par); sp[i] += clock64() - _t; } while (0)
#define SP_LAP(i) do { long long _t = clock64(); sp[i] += _t - sp_t; sp_t = _t; } while (0)
#define SP_RESET() do { sp_t = clock64(); } while (0)
#define SP_DUMP(base) do { if (blockIdx.x == 0) for (int _i = 0; _i < 8; ++_i) g_sprof[(base) + _i] = (unsigned long long)sp[_i]; } while (0)
#define SP_MARK(v) long long v = clock64()
#define SP_ADD(i, v) do { sp[i] += clock64() - v; } while (0)
#else
#define SP_DECL do { } while (0)
#define SP_WAIT(i, bar, par) mbar_wait(bar, par)
#define SP_WAITL(i, bar, par) mbar_wait_long(bar, par)
#define SP_LAP(i) do { } while (0)
#define SP_RESET() do { } while (0)
#define SP_DUMP(base) do { } while (0)
#define SP_MARK(v) do { } while (0)
#define SP_ADD(i, v) do { } while (0)
#endif

constexpr int kSM = 128;                 // latents per row tile
// Scan / apply warps: SP per TMEM lane quadrant, each taking 1/SP of the columns of every code tile ("scan part").
//   SP = 2 (12 warps at 168 registers): d > 128, where a latent's row needs two 16-byte chunks per lane in the apply phase;
//   SP = 4 (20 warps; setmaxnreg: 64 registers for the producer / issuer / converter warpgroup, 104 for the others): d <= 128.
//   The scan is bound by the latency of tcgen05.ld (one load per warp in flight, ~350 clocks per 32-score chunk with two
//   warps per scheduler), the apply phase by L2 round trips with a handful of latents in flight per warp: twice the warps
//   is twice the loads and latents in flight.
__host__ __device__ constexpr int stream_threads(int sp) { return 128 + 128 * sp; }
constexpr int kSCvtWarps = 2;            // converter warps (64 latents each)
constexpr int kSMaxParts = 4;
constexpr int kSCand = 8;                // candidate slots per latent and scan part (the running minimum is seeded: lists stay short)
constexpr int kSXBlock = 8192;           // bytes of one staged x block (converter input, bulk-copied)
constexpr int kSXMaxDepth = 4;           // staged blocks per converter warp, at most
constexpr int kSBrowRing = 4;            // row-bound buffers (converter runs up to 2 tiles ahead of the scan)
constexpr int kSMaxStages = 12;
constexpr int kSOvf = 32;                // spill entries per quadrant and row tile (beyond that: exhaustive scan)
constexpr int kSScratch = kSMaxParts * kSCand + kSOvf;   // merged candidate list of one latent (general resolution path)

struct StreamPlan {
    int stages, stage_bytes, a_bytes;
    int a, b, xs, aone, brow, cs, cc, drop, mfin, mshare, ncnt, recc, recn, ovf, scratch, red, misc, bars, tmem, total;
};
__host__ __device__ inline StreamPlan make_stream_plan(int dp, int nt, int stages, int cg = 1, int sp = 2, int xdepth = 0) {
    StreamPlan u;
    u.stages = stages;
    u.stage_bytes = (nt / cg) * 128;     // this CTA's share of a code slab: NT/cg codes x 64 bf16
    u.a_bytes = kSM * dp * 2;
    int o = 0;
    u.a = o;     o += 2 * u.a_bytes;
    u.b = o;     o += stages * u.stage_bytes;
    u.xs = o;    o += kSCvtWarps * xdepth * kSXBlock;   // staged x blocks: [converter warp][xdepth]
    u.aone = o;  o += kSM * 32;          // constant A block of the |e|^2 step: 128 rows x 16 bf16, SWIZZLE_32B (256-byte aligned)
    u.brow = o;  o += kSBrowRing * kSM * 8;     // per row: |x|^2, |x - bf16(x)|^2
    u.cs = o;    o += sp * kSCand * kSM * 4;
    u.cc = o;    o += sp * kSCand * kSM * 4;
    u.drop = o;  o += 2 * sp * kSM * 4;      // [part][spillmin | dropmin][latent]
    u.mfin = o;  o += sp * kSM * 4;
    u.mshare = o; o += sp * kSM * 4;     // running minimum of every scan part, read by the other parts of the quadrant
    u.ncnt = o;  o += sp * kSM * 4;
    // merged record of every latent (per set): up to four candidate codes and how many (0: general path).  d <= 128 only:
    // at d = 256 the 5 KB are the difference between three and four code-slab stages (8.3 vs 7.3 ms at 16384 x 256)
    u.recc = o;  o += dp <= 128 ? 2 * kSM * 16 : 0;
    u.recn = o;  o += dp <= 128 ? 2 * kSM * 4 : 0;
    u.ovf = o;   o += 2 * 4 * kSOvf * 12;       // [row-tile parity][quadrant]: rows | scores | codes
    u.scratch = o; o += kSScratch * 4 * sp * 4;   // per scan warp
    u.red = o;   o += 32 * 8;
    u.misc = o;  o += 16 * 4;
    u.bars = o;  o += (2 * kSMaxStages + 4 + 8 + 2 * kSBrowRing + kSCvtWarps * kSXMaxDepth) * 8;
    u.tmem = o;  o += 16;
    u.total = o;
    return u;
}

namespace sm100 {
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    // kind::f16: c_format [4,6) = 1 (fp32), a_format [7,10) = 1 (bf16), b_format [10,13) = 1 (bf16), K-major both
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32-bit, 32 consecutive columns per thread.
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
#ifdef TVQ_ABL_NOLD      // ablation build: no TMEM read at all (timing only)
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = taddr + i;
    return;
#endif
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 1-D bulk copy global -> shared (multiple of 16 bytes), completion (bytes) on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// (d0, d1) = (a0, a1) * (b0, b1) + (c0, c1) in one FFMA2
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    uint64_t a, b, c, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float m;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
    return m;
}
// ---- CTA pair (cta_group::2): the even CTA of a 2-CTA cluster issues the MMAs for both
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same object in the EVEN CTA of the pair
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release at cluster scope) on a barrier of the pair's even CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// The same for a hand-off that publishes NO generic-proxy data (a score slot going back to the MMA issuer: the only
// thing ordered is this warp's tcgen05.ld, already complete and fenced by tcgen05.fence::before_thread_sync): the default
// arrive, as CUTLASS's ClusterBarrier::arrive(cta_id) issues it.  The cluster-scope release above made every scan warp
// wait for its candidate-list stores to drain once per code tile (9 % of the kernel's stall samples at 4096 x 128).
__device__ __forceinline__ void mbar_arrive_leader_nodata(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {   // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are counted on the EVEN CTA's barrier (both CTAs of the pair issue it)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_32B: rows of 32 bytes (16 bf16 = one UMMA K step), 8-row
// groups 256 bytes apart — what a TMA box of {16 bf16, rows} with CU_TENSOR_MAP_SWIZZLE_32B writes.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(256 >> 4) << 32;           // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)6 << 61;                    // SWIZZLE_32B
    return d;
}
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
}  // namespace sm100

// Bound on |(s_a - s_b) - (d_a - d_b) / 2| for one latent, s = the tensor-core half-scores |e|^2 / 2 - x^.e^ (x^, e^ the
// bf16 operands; fp32 accumulation incl. the |e|^2 / 2 step), d = the canonical distances.  The operand term uses the
// ACTUAL rounding errors instead of the worst case 2^-9 per operand: x.e - x^.e^ = (x - x^).e + x^.(e - e^), so
// |.| <= |x - x^| max|e| + (|x| + |x - x^|) max|e - e^|, twice that on a score difference.  |x - x^| is measured by the
// converter for every latent, max|e - e^| by the preparation kernel: ~0.4 * 2^-9 relative each on real data — a bound
// 2.5 times tighter than the worst-case form, as rigorous.  Second term: fp32 roundings (accumulation, canonical rule).
__device__ __forceinline__ float stream_bound(const float xn, const float dxn, const float emax, const float demax, const float sum) {
    return fmaf(2.002f, fmaf(dxn, emax, (xn + dxn) * demax), 1.5e-6f * sum * sum);
}

// Spill buffer of one TMEM lane quadrant for the current row tile (shared memory).
struct OvfBuf {
    int* n;        // entries used (may run past kSOvf: then entries were lost)
    int* row;      // [kSOvf] latent (row within the tile)
    float* s;      // [kSOvf] bf16 score
    int* c;        // [kSOvf] code
};

// Scan state of one latent in one scan half (registers of the owning thread).
struct ScanState {
    float m;          // running minimum of the bf16 scores
    int cnt;          // entries in the candidate list
};

// Candidate list of one latent and one scan half: kSCand slots, kSM words apart so that the 32 latents
// of a warp never collide on a bank.  Called when fewer than 4 slots are free (a group of 4 scores is about
// to be appended): compact the list against the current threshold (which only ever decreases); while it
// still holds more than kSCand - 4 entries move the WORST one to the quadrant's spill buffer.  Only if that
// is full too is an entry really lost.  sd[0] / sd[kSM]: smallest spilled / lost score of this latent and
// half (shared memory).  Returns the new count (<= kSCand - 4).
__device__ __noinline__ int cand_make_room(const float thr, float* ls, int* lc, float* sd, const OvfBuf ob, const int trow,
                                           const int cnt) {
    int kept = 0;
    for (int i = 0; i < cnt; ++i) {
        const float v = ls[i * kSM];
        const int c = lc[i * kSM];
        if (v <= thr) { ls[kept * kSM] = v; lc[kept * kSM] = c; ++kept; }
    }
    while (kept > kSCand - 4) {
        int imax = 0;
        float vmax = ls[0];
        for (int i = 1; i < kept; ++i) {
            const float v = ls[i * kSM];
            if (v > vmax) { vmax = v; imax = i; }
        }
        const int pos = atomicAdd(ob.n, 1);
        if (pos < kSOvf) { ob.row[pos] = trow; ob.s[pos] = vmax; ob.c[pos] = lc[imax * kSM]; sd[0] = fminf(sd[0], vmax); }
        else sd[kSM] = fminf(sd[kSM], vmax);
        --kept;
        ls[imax * kSM] = ls[kept * kSM];
        lc[imax * kSM] = lc[kept * kSM];
    }
    return kept;
}

// One chunk of 32 scores of ONE latent (this thread's TMEM lane).  The accumulator already IS the score (halved): the
// GEMM runs on the negated codebook and carries |e|^2 / 2 as one extra K step, s = |e|^2 / 2 - x.e (see the kernel).
__device__ __forceinline__ void scan_chunk(const uint32_t (&r)[32], const int code0, const float brow,
                                           ScanState& st, float* ls, int* lc, float* sd, const OvfBuf& ob, const int trow) {
    using namespace sm100;
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(r[i]);
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fminf(fmin3(s[4 * i], s[4 * i + 1], s[4 * i + 2]), s[4 * i + 3]);
    const float cmin = fmin3(fmin3(g[0], g[1], g[2]), fmin3(g[3], g[4], g[5]), fminf(g[6], g[7]));
    st.m = fminf(st.m, cmin);
    const float thr = st.m + brow;
#ifdef TVQ_ABL_NOSLOW
    if (false)
#endif
    if (cmin <= thr) {
        // which groups of 4 hold a score inside the threshold (straight-line: 8 compares), then one indexed
        // branch per hit group instead of 8 + 4 sequential tests; the appends themselves are call-free
        unsigned hm = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) hm |= (g[i] <= thr) ? (1u << i) : 0u;
        while (hm) {
            const int gi = __ffs((int)hm) - 1;
            hm &= hm - 1u;
            if (st.cnt > kSCand - 4) st.cnt = cand_make_room(thr, ls, lc, sd, ob, trow, st.cnt);
            float a, b, c, d;
            switch (gi) {
                case 0: a = s[0]; b = s[1]; c = s[2]; d = s[3]; break;
                case 1: a = s[4]; b = s[5]; c = s[6]; d = s[7]; break;
                case 2: a = s[8]; b = s[9]; c = s[10]; d = s[11]; break;
                case 3: a = s[12]; b = s[13]; c = s[14]; d = s[15]; break;
                case 4: a = s[16]; b = s[17]; c = s[18]; d = s[19]; break;
                case 5: a = s[20]; b = s[21]; c = s[22]; d = s[23]; break;
                case 6: a = s[24]; b = s[25]; c = s[26]; d = s[27]; break;
                default: a = s[28]; b = s[29]; c = s[30]; d = s[31]; break;
            }
            const int cb = code0 + 4 * gi;
            if (a <= thr) { ls[st.cnt * kSM] = a; lc[st.cnt * kSM] = cb; ++st.cnt; }
            if (b <= thr) { ls[st.cnt * kSM] = b; lc[st.cnt * kSM] = cb + 1; ++st.cnt; }
            if (c <= thr) { ls[st.cnt * kSM] = c; lc[st.cnt * kSM] = cb + 2; ++st.cnt; }
            if (d <= thr) { ls[st.cnt * kSM] = d; lc[st.cnt * kSM] = cb + 3; ++st.cnt; }
        }
    }
}

// Minimum of one chunk of 32 scores (the seed pass: no candidates are recorded).
__device__ __forceinline__ float chunk_min(const uint32_t (&r)[32]) {
    using namespace sm100;
    float g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        g[i] = fminf(fmin3(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2])), __uint_as_float(r[4 * i + 3]));
    return fmin3(fmin3(g[0], g[1], g[2]), fmin3(g[3], g[4], g[5]), fminf(g[6], g[7]));
}

// Straight-through output, commitment-loss partial, EMA statistics and q store of ONE latent
// (all 32 lanes; lane l holds 16-byte chunks l and l + 32 of x and of the chosen code word).
template <int NV, bool TRAIN>
__device__ __forceinline__ void apply_row(const FwdParams& p, float* esum, const float4 xa, const float4 xb, const float4 wa,
                                          const float4 wb, const int code, const int64_t grow, const bool h0, const bool h1,
                                          const int lane, float& loss) {
    float4 oa = wa, ob = wb;
    if (TRAIN) {
        // x + (e - x): two rounded fp32 ops, never contracted; the loss is taken on that rounded tensor,
        // as F.mse_loss(quantize.detach(), x) does
        oa.x = __fadd_rn(xa.x, __fsub_rn(wa.x, xa.x));
        oa.y = __fadd_rn(xa.y, __fsub_rn(wa.y, xa.y));
        oa.z = __fadd_rn(xa.z, __fsub_rn(wa.z, xa.z));
        oa.w = __fadd_rn(xa.w, __fsub_rn(wa.w, xa.w));
        const float dx = __fsub_rn(oa.x, xa.x), dy = __fsub_rn(oa.y, xa.y), dz = __fsub_rn(oa.z, xa.z), dw = __fsub_rn(oa.w, xa.w);
        loss = fmaf(dx, dx, loss); loss = fmaf(dy, dy, loss); loss = fmaf(dz, dz, loss); loss = fmaf(dw, dw, loss);
        if (NV > 1) {
            ob.x = __fadd_rn(xb.x, __fsub_rn(wb.x, xb.x));
            ob.y = __fadd_rn(xb.y, __fsub_rn(wb.y, xb.y));
            ob.z = __fadd_rn(xb.z, __fsub_rn(wb.z, xb.z));
            ob.w = __fadd_rn(xb.w, __fsub_rn(wb.w, xb.w));
            const float ex = __fsub_rn(ob.x, xb.x), ey = __fsub_rn(ob.y, xb.y), ez = __fsub_rn(ob.z, xb.z), ew = __fsub_rn(ob.w, xb.w);
            loss = fmaf(ex, ex, loss); loss = fmaf(ey, ey, loss); loss = fmaf(ez, ez, loss); loss = fmaf(ew, ew, loss);
        }
        float* es_row = esum + (size_t)code * p.d;
        if (h0) red_add_v4(es_row + 4 * lane, xa);
        if (h1) red_add_v4(es_row + 4 * (lane + 32), xb);
    }
    if (p.q != nullptr) {
        float* qr = p.q + (size_t)grow * p.d;
        if (h0) st_stream_v4(qr + 4 * lane, oa);
        if (h1) st_stream_v4(qr + 4 * (lane + 32), ob);
    }
}

// fp32 dot of the lane's x chunks with code word `code`, partial (before the butterfly)
template <int NV>
__device__ __forceinline__ float dot32_part(const float4 x0, const float4 x1, const float* cb, const int code, const int d,
                                            const bool h0, const bool h1, const int lane) {
    const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * d);
    float dd = 0.f;
    if (h0) { const float4 e = __ldg(er + lane); dd = fmaf(x0.x, e.x, fmaf(x0.y, e.y, fmaf(x0.z, e.z, x0.w * e.w))); }
    if (NV > 1 && h1) { const float4 e = __ldg(er + lane + 32); dd = fmaf(x1.x, e.x, fmaf(x1.y, e.y, fmaf(x1.z, e.z, fmaf(x1.w, e.w, dd)))); }
    return dd;
}

// General (rare) resolution of one latent: from a merged candidate list in shared memory (nc > 0) or
// from ALL codes (nc <= 0: entries were lost or the scores were non-finite).  fp32 re-score four codes
// at a time, then the canonical fp64 rule over the same set if the fp32 scores cannot separate the two
// best.  All 32 lanes work on the one latent; lane l holds the 16-byte chunks l (x0) and l + 32 (x1) of x.
// Returns the canonical arg-min, | 1 << 30 if the fp64 level was needed.
template <int NV>
__device__ __noinline__ int resolve_stream(const float4 x0, const float4 x1, const int nc, const int* list, const float* cb,
                                           const float* e2g, const int k, const int d, const float emax, const int lane) {
    const int nchunk = d >> 2;
    const bool h0 = lane < nchunk, h1 = NV > 1 && lane + 32 < nchunk;
    const float INF = __int_as_float(0x7f800000);
    float ss = fmaf(x0.x, x0.x, fmaf(x0.y, x0.y, fmaf(x0.z, x0.z, x0.w * x0.w)));
    if (NV > 1) ss = fmaf(x1.x, x1.x, fmaf(x1.y, x1.y, fmaf(x1.z, x1.z, fmaf(x1.w, x1.w, ss))));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float bnd = fmaf(sqrtf(ss), 1.0001f, emax);
    const float thr32 = 1.6e-6f * bnd * bnd;
    const int count = nc > 0 ? nc : k;
    float m1 = INF, m2 = INF;
    int i1 = 0;
    for (int j0 = 0; j0 < count; j0 += 4) {
        int code[4];
        float dd[4], e2v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = j0 + t < count ? j0 + t : count - 1;
            int c = nc > 0 ? list[j] : j;
            c = (c >= 0 && c < k) ? c : 0;
            code[t] = c;
            dd[t] = dot32_part<NV>(x0, x1, cb, c, d, h0, h1, lane);
            e2v[t] = __ldg(e2g + c);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int t = 0; t < 4; ++t) dd[t] += __shfl_xor_sync(0xffffffffu, dd[t], off);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (j0 + t >= count) break;
            const float sv = fmaf(-2.f, dd[t], e2v[t]);
            if (sv < m1) { m2 = m1; m1 = sv; i1 = code[t]; }
            else if (sv < m2) m2 = sv;
        }
    }
    if (m2 - m1 > thr32) return i1;                       // NaN compares false: falls through to fp64
    double pp = 0.0;
    if (h0) pp = dot4(pp, x0, x0);
    if (h1) pp = dot4(pp, x1, x1);
    const float x2 = __double2float_rn(butterfly_sum(pp));
    float best = INF;
    int arg = 0x7fffffff;
    for (int j = 0; j < count; ++j) {
        int code = nc > 0 ? list[j] : j;
        code = (code >= 0 && code < k) ? code : 0;
        const float4* er = reinterpret_cast<const float4*>(cb + (size_t)code * d);
        double sd = 0.0;
        if (h0) sd = dot4(sd, x0, __ldg(er + lane));
        if (h1) sd = dot4(sd, x1, __ldg(er + lane + 32));
        const float dk = canon_score(x2, butterfly_sum(sd), __ldg(e2g + code));
        if (dk < best || (dk == best && code < arg)) { best = dk; arg = code; }
    }
    if (arg == 0x7fffffff) arg = 0;                       // non-finite row: any valid code
    return arg | (1 << 30);
}

// Merge the candidates of one latent into `list` (this warp's scratch): every part's entries (cnt[part] of them, in part
// order), and the quadrant's spilled entries of this latent that lie inside the final threshold.  Warp-cooperative.
template <int SP>
__device__ __forceinline__ int gather_cands(int* list, const int (&cnts)[SP], const int* lc_row, const OvfBuf ob,
                                            const int lrow, const float thr, const int lane) {
    int cnt = 0;
#pragma unroll
    for (int pt = 0; pt < SP; ++pt) {
        if (lane < cnts[pt]) list[cnt + lane] = lc_row[(pt * kSCand + lane) * kSM];
        cnt += cnts[pt];
    }
    const int on = min(*ob.n, kSOvf);
    for (int base = 0; base < on; base += 32) {
        const int e = base + lane;
        const bool match = e < on && ob.row[e] == lrow && ob.s[e] <= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, match);
        if (match) list[cnt + __popc(bal & lanemask_lt())] = ob.c[e];
        cnt += __popc(bal);
    }
    __syncwarp();
    return cnt;
}

// CG = 1: one CTA per SM on its own.  CG = 2: CTA pairs (2-CTA clusters, tcgen05 cta_group::2): each CTA keeps its
// own 128 latents and everything that belongs to them (converters, scan warps, lists, TMEM scores), but only HALF of
// every code slab: the even CTA issues one M = 256 MMA for both, which reads the A rows and the B half of each
// CTA from that CTA's shared memory.  Per SM this halves the TMA fill and the L2 traffic of the code stream and
// takes a third off the shared-memory operand reads.
// SETS = 2 (d <= 128): the scan / apply warps form TWO sets of 4 x SP warps; set s takes the row tiles with it % 2 == s
// and owns half of the tensor-memory score slots.  While one set resolves and applies its row tile (L2 round trips, no
// tensor-memory reads) the MMA issuer already fills the other set's slots and that set scans: the apply phase no longer
// stalls the tensor pipe (it used to: 14 k of 64 k clocks per row tile at 4096 x 128).  Storage is indexed by set * SP + part.
template <int DP, int NT, bool TRAIN, int CG, int SP, int SETS>
__global__ void __launch_bounds__(stream_threads(SP * SETS), 1) fwd_stream_kernel(const __grid_constant__ CUtensorMap tmap_cb,
                                                                 const __grid_constant__ CUtensorMap tmap_e2, const FwdParams p,
                                                                 const int stages, const __nv_bfloat16* __restrict__ e2h, const int xdepth) {
    using namespace sm100;
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int KSLABS = DP / 64;             // 64-column (128-byte) bf16 slabs per row
    constexpr int SLOTS = 512 / NT;             // tensor-memory score slots
    constexpr int SLOTS_S = SLOTS / SETS;       // ... of one set
    constexpr int NCH = NT / (32 * SP);         // 32-score chunks per code tile AND scan part
    constexpr int kSThreads = stream_threads(SP * SETS);
    constexpr int kSetWarps = 4 * SP;           // scan / apply warps of one set
    constexpr int LPW = 32 / SP;                // latents per warp in the apply phase
    constexpr bool WG_REGS = SP * SETS == 4;    // 20 warps: registers shared out by warpgroup (setmaxnreg)
    static_assert(SP == 2 || SP == 4, "two or four scan parts per quadrant");
    static_assert(SETS == 1 || SETS == 2, "one or two scan / apply sets");
    static_assert(SLOTS_S >= 2, "at least two score slots per set");
    static_assert(NCH >= 1, "at least one chunk per part");
    constexpr int F = DP / 4;                   // 16-byte fp32 chunks per padded row
    constexpr int NV = DP > 128 ? 2 : 1;        // 16-byte chunks per lane in the warp-per-latent phases
    constexpr int A_SLAB = kSM * 128;           // bytes of one A slab (128 rows x 128 bytes)
    static_assert(DP == 64 || DP == 128 || DP == 256, "d padded to 64, 128 or 256");
    static_assert(NT == 128 || NT == 256, "code tile of 128 or 256");

    const StreamPlan pl = make_stream_plan(DP, NT, stages, CG, SP * SETS, xdepth);
    float2* brow_ring = reinterpret_cast<float2*>(smem + pl.brow);
    float* cand_s = reinterpret_cast<float*>(smem + pl.cs);     // [part][slot][latent]
    int* cand_c = reinterpret_cast<int*>(smem + pl.cc);
    float* mfin = reinterpret_cast<float*>(smem + pl.mfin);     // [half][latent] minimum seen by each scan half
    volatile float* mshare = reinterpret_cast<volatile float*>(smem + pl.mshare);   // [part][latent] running minimum, live
    int* ncnt = reinterpret_cast<int*>(smem + pl.ncnt);         // [half][latent] final candidates per half (-1: incomplete)
    int4* recc_base = reinterpret_cast<int4*>(smem + pl.recc);
    int* recn_base = reinterpret_cast<int*>(smem + pl.recn);
    int* ovf_base = reinterpret_cast<int*>(smem + pl.ovf);      // [parity][quadrant][rows | scores | codes][kSOvf]
    int* scratch_base = reinterpret_cast<int*>(smem + pl.scratch);
    double* red = reinterpret_cast<double*>(smem + pl.red);
    int* misc = reinterpret_cast<int*>(smem + pl.misc);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + pl.tmem);
    const uint32_t bar0 = smem_u32(smem + pl.bars);
    const uint32_t bar_bfull = bar0, bar_bempty = bar_bfull + 8 * kSMaxStages;
    const uint32_t bar_afull = bar_bempty + 8 * kSMaxStages, bar_aempty = bar_afull + 16;
    const uint32_t bar_tfull = bar_aempty + 16, bar_tempty = bar_tfull + 32;
    const uint32_t bar_rfull = bar_tempty + 32, bar_rempty = bar_rfull + 8 * kSBrowRing;
    const uint32_t bar_xfull = bar_rempty + 8 * kSBrowRing;   // [converter warp][kSXMaxDepth]
    const uint32_t a_base = smem_u32(smem + pl.a), b_base = smem_u32(smem + pl.b), aone_base = smem_u32(smem + pl.aone);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = p.d >> 2;
    const int n_ct = (p.k + NT - 1) / NT;                 // code tiles
    const int num_tiles = p.num_tiles;                    // row tiles of 128 latents
    // work distribution: unit u = blockIdx.x / CG walks the tile groups; this CTA takes tile CG * group + rank
    // (possibly past the end: such a tile has no valid latents but the CTA still takes part in the pair's MMAs)
    const int crank = CG == 2 ? (int)cluster_ctarank() : 0;
    const int unit0 = (int)blockIdx.x / CG, nunits = (int)gridDim.x / CG;
    const int ngroups = (num_tiles + CG - 1) / CG;

    // ------------------------------------------------------------------ CTA prologue
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    if (tid == 0) {
        for (int s = 0; s < kSMaxStages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_afull + 8 * s, CG * kSCvtWarps); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, CG * kSetWarps); }
        for (int s = 0; s < kSBrowRing; ++s) { mbar_init(bar_rfull + 8 * s, kSCvtWarps); mbar_init(bar_rempty + 8 * s, kSetWarps); }
        for (int s = 0; s < kSCvtWarps * kSXMaxDepth; ++s) mbar_init(bar_xfull + 8 * s, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmap_cb);
        tma_prefetch_desc(&tmap_e2);
        for (int i = 4; i < 12; ++i) misc[i] = 0;           // spill counters [parity][quadrant]
    }
    if (warp == 1) { if (CG == 2) tmem_alloc2(smem_u32(tmem_slot), 512); else tmem_alloc(smem_u32(tmem_slot), 512); }
    // constant A block of the |e|^2 step: every row is (1, 1, 1, 0, ..., 0) in bf16; 16-byte chunk c of row r sits at
    // chunk (c ^ ((r >> 2) & 1)) of the row's 32 bytes (SWIZZLE_32B)
    for (int i = tid; i < kSM * 2; i += kSThreads) {
        const int r = i >> 1, c = i & 1;
        const uint4 v = c == 0 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(smem + pl.aone + r * 32 + ((c ^ ((r >> 2) & 1)) << 4)) = v;
    }
    if (CG == 2) fence_proxy_async_all(); else fence_proxy_async_smem();
    // max |e| (error-bound constant): every CTA scans the |e|^2 table (k floats, L2 resident)
    float emax2 = 0.f, dmax2 = 0.f;
    for (int c = tid; c < p.k; c += kSThreads) {
        emax2 = fmaxf(emax2, __ldg(p.e2 + c));
        dmax2 = fmaxf(dmax2, __bfloat162float(e2h[(size_t)c * 16 + 3]));    // |e_c - bf16(e_c)|^2 (prep_kernel)
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        emax2 = fmaxf(emax2, __shfl_xor_sync(0xffffffffu, emax2, off));
        dmax2 = fmaxf(dmax2, __shfl_xor_sync(0xffffffffu, dmax2, off));
    }
    float* wmax = reinterpret_cast<float*>(red);
    if (lane == 0) { wmax[warp] = emax2; wmax[32 + warp] = dmax2; }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();                      // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    emax2 = 0.f; dmax2 = 0.f;
    for (int w = 0; w < kSThreads / 32; ++w) { emax2 = fmaxf(emax2, wmax[w]); dmax2 = fmaxf(dmax2, wmax[32 + w]); }
    const float emax = sqrtf(emax2) * 1.0001f;
    const float demax = sqrtf(dmax2) * 1.0001f;
    __syncthreads();                                      // wmax (aliases red) is free again

    double loss_d = 0.0;
    unsigned n_resc = 0, n_f64 = 0;

    // 20 warps (SP == 4): the register file is shared out by warpgroup — 96 per thread at launch (61 440 in all: the pool setmaxnreg draws on), 64 for the producer /
    // issuer / converter warpgroup, 104 for the scan / apply warpgroups (setmaxnreg at the head of each branch, so that the
    // register allocator sees it dominate the branch)
    if (warp < 4) {
    if (WG_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ============================================================ producer: code slabs (TMA) + |e|^2 slices
        // per code tile: KSLABS slabs of the (negated) bf16 codebook, then the |e|^2 / 2 slab [NT x 16 bf16] (its table in the
        // workspace is padded with BIG to a multiple of 256 codes, so codes >= k can never be nominated).
        // The whole warp walks the loop (warp-uniform control flow keeps the counters in uniform registers); one elected
        // lane issues.  Stage index and phase are carried, not divided out of a running counter.
        {
            const bool leader = elect_one();
            SP_DECL;
            int s = 0;
            uint32_t ph = 0;
            for (int grp = unit0; grp < ngroups; grp += nunits) {
                for (int ct = 0; ct < n_ct; ++ct) {
#pragma unroll 1
                    for (int j = 0; j <= KSLABS; ++j) {
                        const uint32_t bytes = j < KSLABS ? (uint32_t)(NT * 128) : (uint32_t)(NT * 32);
                        const CUtensorMap* tm = j < KSLABS ? &tmap_cb : &tmap_e2;
                        const int c0 = j < KSLABS ? j * 64 : 0;
                        SP_WAIT(1, bar_bempty + 8 * s, ph ^ 1u);
                        if (leader) {
                            if (CG == 2) {
                                // the even CTA's barrier counts the bytes of both halves; each CTA fetches its half
                                if (crank == 0) mbar_arrive_expect_tx(bar_bfull + 8 * s, bytes);
                                tma_load_2d_pair(b_base + s * pl.stage_bytes, tm, bar_bfull + 8 * s, c0, ct * NT + crank * (NT / 2));
                            } else {
                                mbar_arrive_expect_tx(bar_bfull + 8 * s, bytes);
                                tma_load_2d(b_base + s * pl.stage_bytes, tm, bar_bfull + 8 * s, c0, ct * NT);
                            }
                        }
                        __syncwarp();
                        if (++s == stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
            SP_LAP(7);
            if (lane == 0) SP_DUMP(0);
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer (the even CTA of a pair issues for both)
        if (crank == 0) {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc_bf16(CG * kSM, NT);
            SP_DECL;
            int s = 0, it = 0;
            uint32_t ph = 0;
            int tts0 = 0, tts1 = 0;                           // code tiles issued for each set
            for (int grp = unit0; grp < ngroups; grp += nunits, ++it) {
                const int ab = it & 1;
                const int set = SETS == 2 ? (it & 1) : 0;
                SP_WAIT(0, bar_afull + 8 * ab, ((uint32_t)(it >> 1)) & 1u);
                tc_fence_after();
                const uint32_t a0 = a_base + ab * pl.a_bytes;
                for (int ct = 0; ct < n_ct; ++ct) {
                    const int tt = set == 0 ? tts0++ : tts1++;
                    const int slot = set * SLOTS_S + tt % SLOTS_S;
                    SP_WAIT(1, bar_tempty + 8 * slot, (((uint32_t)(tt / SLOTS_S)) & 1u) ^ 1u);
                    tc_fence_after();
#pragma unroll 1
                    for (int j = 0; j <= KSLABS; ++j) {
                        SP_WAIT(2, bar_bfull + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t b0 = b_base + s * pl.stage_bytes;
                        if (leader) {
                            if (j < KSLABS) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
#ifdef TVQ_ABL_NOMMA     // ablation build: no MMA issued (timing of the scan without tensor-pipe / operand traffic)
                                    continue;
#endif
                                    if (CG == 2)
                                        umma_bf16_pair(tmem_base + slot * NT, umma_desc_sw128(a0 + j * A_SLAB + kk * 32),
                                                       umma_desc_sw128(b0 + kk * 32), idesc, (j | kk) != 0);
                                    else
                                        umma_bf16(tmem_base + slot * NT, umma_desc_sw128(a0 + j * A_SLAB + kk * 32),
                                                  umma_desc_sw128(b0 + kk * 32), idesc, (j | kk) != 0);
                                }
                            } else {
                                // + |e|^2 / 2: ONE more K step, constant A rows (1, 1, 1, 0...) x the three bf16 pieces of |e_c|^2 / 2
                                if (CG == 2) umma_bf16_pair(tmem_base + slot * NT, umma_desc_sw32(aone_base), umma_desc_sw32(b0), idesc, 1u);
                                else umma_bf16(tmem_base + slot * NT, umma_desc_sw32(aone_base), umma_desc_sw32(b0), idesc, 1u);
                            }
                            // stage free (in both CTAs of a pair) once these MMAs have read it
                            if (CG == 2) umma_commit_pair(bar_bempty + 8 * s); else umma_commit(bar_bempty + 8 * s);
                            // scores of this code tile complete
                            if (j == KSLABS) { if (CG == 2) umma_commit_pair(bar_tfull + 8 * slot); else umma_commit(bar_tfull + 8 * slot); }
                        }
                        __syncwarp();
                        if (++s == stages) { s = 0; ph ^= 1u; }
                    }
                }
                // A buffer free once every MMA of the row tile is done
                if (leader) { if (CG == 2) umma_commit_pair(bar_aempty + 8 * ab); else umma_commit(bar_aempty + 8 * ab); }
                __syncwarp();
            }
            SP_LAP(7);
            if (lane == 0) SP_DUMP(8);
        }
    } else if (warp < 2 + kSCvtWarps) {
        // ============================================================ converters: x fp32 -> bf16 A operand, row norms
        // Each converter warp is ONE serial instruction stream, so what matters is how few instructions a 16-byte chunk costs
        // (the first version spent ~100, 24 k clocks per row tile at d = 64: the floor of every k <= 1024 shape):
        //   * d <= 128: the x rows arrive through shared memory — 8 KB blocks (512 / F rows) bulk-copied by this warp xdepth - 1
        //     blocks ahead; a lane owns a ROW (or half of one), so per chunk it is one LDS.128, two cvt.rn.bf16x2, one STS.64
        //     into the swizzled A tile and the two sums of squares as per-lane accumulations; all index arithmetic is hoisted;
        //   * the square roots and the bound itself are left to the scan threads (one per latent): the converter only
        //     publishes |x|^2 and |x - bf16(x)|^2 per row.
        const int cw = warp - 2;
        constexpr bool STAGED = DP <= 128;
        constexpr int U = STAGED ? 4 : 8;                     // 16-byte chunks per lane in flight
        constexpr int XRB = STAGED ? 512 / F : 32;            // rows per block
        constexpr int BPT = 64 / XRB;                         // blocks per row tile and converter warp
        const uint32_t xs_base = smem_u32(smem + pl.xs) + (uint32_t)(cw * xdepth) * kSXBlock;
        const uint32_t xbar0 = bar_xfull + 8 * (cw * kSXMaxDepth);
        auto issue_block = [&](const int g) {                 // lane 0: bulk copy of block g of this warp's block sequence
            const int it2 = g / BPT, blk2 = g - it2 * BPT;
            const int64_t grp2 = (int64_t)unit0 + (int64_t)it2 * nunits;
            if (grp2 >= ngroups) return;
            const int64_t r0 = (CG * grp2 + crank) * kSM + cw * 64 + blk2 * XRB;
            int64_t rows = p.n - r0;
            rows = rows < XRB ? rows : XRB;
            if (rows <= 0) return;
            const uint32_t bytes = (uint32_t)rows * (uint32_t)p.d * 4u;
            const int st = g % xdepth;
            mbar_arrive_expect_tx(xbar0 + 8 * st, bytes);
            bulk_load_1d(xs_base + (uint32_t)st * kSXBlock, p.x + (size_t)r0 * p.d, bytes, xbar0 + 8 * st);
        };
        if (STAGED && lane == 0)
            for (int g = 0; g < xdepth - 1; ++g) issue_block(g);
        // per-lane constants of the staged path: lane -> (row brw of a block, part sub of the row), see the loop
        constexpr int LPR = STAGED ? 32 / XRB : 1;            // lanes per row
        static_assert(!STAGED || (F / LPR == 16 && (LPR == 1 || LPR == 2)), "a lane converts 16 chunks of one row");
        const int brw = lane / LPR;
        const uint32_t sub = (uint32_t)(lane % LPR);
        const uint32_t rx = (uint32_t)brw & 7u;               // row & 7 of the A tile (blocks start at multiples of 16 rows)
        const int rot = brw + 8 * (int)sub;
        const uint32_t rowbytes = (uint32_t)p.d * 4u;
        SP_DECL;
        int it = 0, gblk = 0;
        for (int grp = unit0; grp < ngroups; grp += nunits, ++it) {
            const int tile = CG * grp + crank;
            const int ab = it & 1;
            const int64_t row0 = (int64_t)tile * kSM;
            if (cw == 0 && lane == 0) {                       // L2 prefetch of the tile two iterations ahead
                const int64_t t2 = (int64_t)tile + 2 * (int64_t)gridDim.x;
                if (t2 < num_tiles) {
                    const int64_t r2 = t2 * kSM;
                    const int64_t rows = (p.n - r2) < kSM ? (p.n - r2) : kSM;
                    prefetch_l2_bulk(p.x + (size_t)r2 * p.d, (uint32_t)(rows * p.d * 4));
                }
            }
            const int rs = it & (kSBrowRing - 1);
            SP_WAITL(0, bar_aempty + 8 * ab, (((uint32_t)(it >> 1)) & 1u) ^ 1u);
            SP_WAIT(1, bar_rempty + 8 * rs, (((uint32_t)(it / kSBrowRing)) & 1u) ^ 1u);
            SP_RESET();
            const uint32_t a0 = a_base + ab * pl.a_bytes;
            float2* brow = brow_ring + rs * kSM;
#pragma unroll 1
            for (int blk = 0; blk < BPT; ++blk, ++gblk) {
            const int rowbase = cw * 64 + blk * XRB;           // first latent of the block within the tile
            if constexpr (STAGED) {
                const int64_t left = p.n - (row0 + rowbase);
                const int nvalid = left < XRB ? (int)(left > 0 ? left : 0) : XRB;    // rows of this block that exist
                __syncwarp();                                 // every lane has read the block whose buffer is refilled now
                SP_MARK(sp_c0);
                if (lane == 0) { fence_proxy_async_smem(); issue_block(gblk + xdepth - 1); }
                SP_ADD(3, sp_c0);
                SP_MARK(sp_c1);
                if (nvalid > 0) mbar_wait(xbar0 + 8 * (gblk % xdepth), ((uint32_t)(gblk / xdepth)) & 1u);
                SP_ADD(4, sp_c1);
                // lane <-> ROW: a lane converts 16 chunks of one row of the block (d <= 64: 32 rows of 16 chunks, one lane per
                // row; d <= 128: 16 rows of 32 chunks, two lanes per row), so the two sums of squares of a row are plain
                // per-lane accumulations — no shuffle butterfly, no select network (the lane <-> chunk form spent half of its
                // ~42 instructions per chunk and all of its dependent latency there: 20 k clocks per row tile at d = 128, the
                // floor of every k <= 1024 shape at d >= 128).  The chunk order is rotated by the row (+ 8 for the second lane
                // of a row): LDS.128 at most 2-way and STS.64 not at all bank-conflicted, although the rows are 256 / 512 B apart.
                const bool rvalid = brw < nvalid;
                const uint32_t src = xs_base + (uint32_t)(gblk % xdepth) * kSXBlock + (uint32_t)brw * rowbytes + (uint32_t)sub * 256u;
                const uint32_t dst = a0 + (uint32_t)sub * A_SLAB + (uint32_t)(rowbase + brw) * 128u;
                float s1 = 0.f, s2 = 0.f;                     // |x|^2 and |x - bf16(x)|^2 of this lane's part of the row
#pragma unroll 1
                for (int j0 = 0; j0 < 16; j0 += U) {
                    float4 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t ccl = (uint32_t)(j0 + u + rot) & 15u;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rvalid && (int)(sub * 16u + ccl) < nchunk) v[u] = lds_v4(src + ccl * 16u);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t ccl = (uint32_t)(j0 + u + rot) & 15u;
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y);
                        const __nv_bfloat162 hi = __floats2bfloat162_rn(v[u].z, v[u].w);
                        const uint32_t wlo = *reinterpret_cast<const uint32_t*>(&lo), whi = *reinterpret_cast<const uint32_t*>(&hi);
                        // element column 4 * (16 sub + ccl): slab sub, 16-byte chunk ccl / 2 (XOR row & 7), half ccl & 1
                        sts_v2(dst + ((((ccl >> 1) ^ rx)) << 4) + ((ccl & 1u) << 3), wlo, whi);
                        s1 += fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
                        // |x - bf16(x)|^2 of the chunk (a bf16 is the upper half of its fp32; the differences are exact)
                        const float ex = v[u].x - __uint_as_float(wlo << 16), ey = v[u].y - __uint_as_float(wlo & 0xffff0000u);
                        const float ez = v[u].z - __uint_as_float(whi << 16), ew = v[u].w - __uint_as_float(whi & 0xffff0000u);
                        s2 += fmaf(ex, ex, fmaf(ey, ey, fmaf(ez, ez, ew * ew)));
                    }
                }
                if (LPR == 2) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                }
                if (sub == 0) brow[rowbase + brw] = make_float2(s1, s2);
            } else {
                // d = 256 (no room for staged blocks): global loads, eight 16-byte chunks per lane in flight = four rows per
                // iteration (a row is two warp-wide accesses); the same hoisted per-lane indexing and folded reduction
                static_assert(STAGED || (F == 64 && U == 8), "the global-load converter path serves d = 256 only");
                const int64_t left = p.n - (row0 + rowbase);
                const int nvalid = left < XRB ? (int)(left > 0 ? left : 0) : XRB;    // rows of this block that exist
                const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)(row0 + rowbase) * p.d) + lane;
                const uint32_t dst = a0 + (uint32_t)rowbase * 128u + ((uint32_t)(lane & 1) << 3);
                const uint32_t slab_l = (uint32_t)(lane >> 4) * A_SLAB, chunk_g = (uint32_t)(lane & 15) >> 1;
                const bool cv0 = lane < nchunk, cv1 = lane + 32 < nchunk;
#pragma unroll 1
                for (int i0 = 0; i0 < XRB; i0 += 4) {          // four rows
                    float4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = i0 + (u >> 1);
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r < nvalid && ((u & 1) ? cv1 : cv0)) v[u] = __ldg(src + (size_t)r * (size_t)(p.d >> 2) + 32 * (u & 1));
                    }
                    float vals[8];                            // [r]: |x|^2 of row i0 + r, [4 + r]: |x - bf16(x)|^2
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = i0 + (u >> 1);
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y);
                        const __nv_bfloat162 hi = __floats2bfloat162_rn(v[u].z, v[u].w);
                        const uint32_t wlo = *reinterpret_cast<const uint32_t*>(&lo), whi = *reinterpret_cast<const uint32_t*>(&hi);
                        // chunk c4 = lane + 32 (u & 1): slab c4 / 16, 16-byte chunk (c4 % 16) / 2 (XOR row & 7), half c4 & 1
                        sts_v2(dst + slab_l + (uint32_t)(2 * (u & 1)) * A_SLAB + (uint32_t)r * 128u + ((chunk_g ^ ((uint32_t)(rowbase + r) & 7u)) << 4),
                               wlo, whi);
                        const float t = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
                        const float ex = v[u].x - __uint_as_float(wlo << 16), ey = v[u].y - __uint_as_float(wlo & 0xffff0000u);
                        const float ez = v[u].z - __uint_as_float(whi << 16), ew = v[u].w - __uint_as_float(whi & 0xffff0000u);
                        const float td = fmaf(ex, ex, fmaf(ey, ey, fmaf(ez, ez, ew * ew)));
                        if (u & 1) { vals[u >> 1] += t; vals[4 + (u >> 1)] += td; }
                        else { vals[u >> 1] = t; vals[4 + (u >> 1)] = td; }
                    }
                    {
                        const bool b0 = (lane & 16) != 0, b1 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float keep = b0 ? vals[i + 4] : vals[i], send = b0 ? vals[i] : vals[i + 4];
                            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float keep = b1 ? vals[i + 2] : vals[i], send = b1 ? vals[i] : vals[i + 2];
                            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
                        {
                            const float keep = b2 ? vals[1] : vals[0], send = b2 ? vals[0] : vals[1];
                            vals[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                        vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], 2);
                        vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], 1);
                        if ((lane & 3) == 0)
                            reinterpret_cast<float*>(brow)[2 * (rowbase + i0 + 2 * (int)b1 + (int)b2) + (int)b0] = vals[0];
                    }
                }
            }
            }
            SP_MARK(sp_c2);
            if (CG == 2) fence_proxy_async_all(); else fence_proxy_async_smem();   // generic-proxy stores -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_leader(bar_afull + 8 * ab); else mbar_arrive(bar_afull + 8 * ab);
                mbar_arrive(bar_rfull + 8 * rs);
            }
            SP_ADD(5, sp_c2);
            SP_LAP(2);
        }
        if (cw == 0 && lane == 0) SP_DUMP(24);
    }
    } else {
        if (WG_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ============================================================ scan + apply warps (one thread per latent)
        const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
        const int set = (warp - 4) / kSetWarps;               // which row tiles (it % SETS) and score slots this warp works on
        const int half = ((warp - 4) >> 2) % SP;              // scan part: which 1/SP of every code tile's columns this warp scans
        const int pbase = set * SP, pidx = pbase + half;      // storage index of the set's first part / of this part
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * (NT / SP));
        const int trow = quad * 32 + lane;                    // this thread's latent within the tile
        float* ls = cand_s + pidx * (kSCand * kSM) + trow;
        int* lc = cand_c + pidx * (kSCand * kSM) + trow;
        float* sd = reinterpret_cast<float*>(smem + pl.drop) + pidx * (2 * kSM) + trow;   // [0]: spillmin, [kSM]: dropmin
        float* thrfin = mfin + pbase * kSM;                   // reused after the merge: final threshold per latent (part 0 slot)
        int* scratch = scratch_base + (warp - 4) * kSScratch;
        const uint32_t qbar = 1u + (uint32_t)(set * 4 + quad);   // named barrier of this set's quadrant (SP warps)
        int4* recc = recc_base + set * kSM;
        int* recn = recn_base + set * kSM;
        float* esum = p.stats + ((p.k + 3) & ~3);
        const bool h0 = lane < nchunk, h1 = NV > 1 && lane + 32 < nchunk;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float INF = __int_as_float(0x7f800000);
        // latents in flight in the apply phase (up to 4 candidates each are resolved in the batched pass).  Two candidates
        // per latent with twice the latents per L2 round trip was measured and is slower (512 x 64: 6.3 vs 4.4 ms): three-
        // and four-candidate latents are too common for the one-at-a-time second pass.
        SP_DECL;
        int et = 0;                                           // code tiles scanned by this set
        for (int it = set, grp = unit0 + set * nunits; grp < ngroups; grp += SETS * nunits, it += SETS) {
            const int tile = CG * grp + crank;
            const int rs = it & (kSBrowRing - 1);
            SP_WAIT(0, bar_rfull + 8 * rs, ((uint32_t)(it / kSBrowRing)) & 1u);     // the row bounds are written
            const float2 nrm = brow_ring[rs * kSM + trow];     // |x|^2 and |x - bf16(x)|^2 of this thread's latent (converter)
            const float xnrm = sqrtf(nrm.x) * 1.0001f;
            const float brow = stream_bound(xnrm, sqrtf(nrm.y) * 1.0001f, emax, demax, xnrm + emax);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rempty + 8 * rs);
            ScanState st;
            st.m = INF; st.cnt = 0;
            sd[0] = INF; sd[kSM] = INF;
            OvfBuf ob;
            {
                int* ob0 = ovf_base + ((it & 1) * 4 + quad) * (3 * kSOvf);
                ob.n = misc + 4 + (it & 1) * 4 + quad;
                ob.row = ob0;
                ob.s = reinterpret_cast<float*>(ob0 + kSOvf);
                ob.c = ob0 + 2 * kSOvf;
            }
            // ---- seed: the minimum over the FIRST code tile (all parts of the quadrant), before anything is recorded.  A
            //      running minimum that starts at +inf makes every early score a "new record": with k codes a latent sets
            //      ~ln k of them, and with 32 latents per warp nearly every chunk took the candidate path.  Seeded, the first
            //      tile records only what lies within the bound of its own minimum, and a later chunk at position n is
            //      entered with probability ~32 / n per latent.
            {
                const int slot = set * SLOTS_S + et % SLOTS_S;
                SP_WAIT(2, bar_tfull + 8 * slot, ((uint32_t)(et / SLOTS_S)) & 1u);
                tc_fence_after();
                SP_RESET();
                const uint32_t taddr = lane_addr + (uint32_t)(slot * NT);
                uint32_t ra[32], rb[32];
                float mp = INF;
#pragma unroll 1
                for (int c = 0; c < NCH; c += 2) {
                    tmem_ld_x32(taddr + (uint32_t)c * 32u, ra);
                    if (c + 1 < NCH) tmem_ld_x32(taddr + (uint32_t)(c + 1) * 32u, rb);
                    tmem_ld_wait();
                    mp = fminf(mp, chunk_min(ra));
                    if (c + 1 < NCH) mp = fminf(mp, chunk_min(rb));
                }
                mshare[pidx * kSM + trow] = mp;
                if (half == 0 && lane == 0) *ob.n = 0;        // the quadrant's spill buffer: nothing is recorded before this barrier
                named_bar_sync(qbar, 32 * SP);
#pragma unroll
                for (int pt = 0; pt < SP; ++pt) st.m = fminf(st.m, mshare[(pbase + pt) * kSM + trow]);
                SP_LAP(3);
            }
            // ---- scan: this warp's part of the columns of every code tile.  The parts of a quadrant share their running
            //      minima (any score of the latent is a valid upper bound of its minimum; a stale value only costs a
            //      candidate that the final filter drops), exchanged once per code tile through shared memory.
            for (int ct = 0; ct < n_ct; ++ct, ++et) {
                const int slot = set * SLOTS_S + et % SLOTS_S;
                SP_WAIT(2, bar_tfull + 8 * slot, ((uint32_t)(et / SLOTS_S)) & 1u);
                tc_fence_after();
                SP_RESET();
                if (ct > 0) {
#pragma unroll
                    for (int pt = 0; pt < SP; ++pt) st.m = fminf(st.m, mshare[(pbase + pt) * kSM + trow]);
                }
                const uint32_t taddr = lane_addr + (uint32_t)(slot * NT);
                const int code0 = ct * NT + half * (NT / SP);
                uint32_t ra[32], rb[32];
                // The score slot goes back to the MMA issuer as soon as this warp's LAST tcgen05.ld of the code tile has
                // completed — before the scores are looked at.  A slot is released by the slowest of the (pair's) scan warps,
                // and some warp takes the candidate path in nearly every code tile: held through the scan, the slot paid for
                // that path every time.
                auto release_slot = [&]() {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_leader_nodata(bar_tempty + 8 * slot); else mbar_arrive(bar_tempty + 8 * slot);
                    }
                };
                tmem_ld_x32(taddr, ra);
                if constexpr (NCH == 1) {
                    tmem_ld_wait();
                    release_slot();
                    scan_chunk(ra, code0, brow, st, ls, lc, sd, ob, trow);
                } else if constexpr (NCH == 2) {
                    tmem_ld_x32(taddr + 32u, rb);
                    tmem_ld_wait();
                    release_slot();
                    scan_chunk(ra, code0, brow, st, ls, lc, sd, ob, trow);
                    scan_chunk(rb, code0 + 32, brow, st, ls, lc, sd, ob, trow);
                } else {
#pragma unroll 1
                    for (int c = 0; c < NCH; c += 2) {      // the next chunk's tcgen05.ld is in flight while this one is scanned
                        tmem_ld_wait();
                        tmem_ld_x32(taddr + (uint32_t)(c + 1) * 32u, rb);
                        scan_chunk(ra, code0 + c * 32, brow, st, ls, lc, sd, ob, trow);
                        tmem_ld_wait();
                        if (c + 2 < NCH) tmem_ld_x32(taddr + (uint32_t)(c + 2) * 32u, ra);
                        else release_slot();
                        scan_chunk(rb, code0 + (c + 1) * 32, brow, st, ls, lc, sd, ob, trow);
                    }
                }
                mshare[pidx * kSM + trow] = st.m;
                SP_LAP(3);
            }
            SP_RESET();
            // ---- merge the two halves of the quadrant: common minimum, then each half filters its own list
            mfin[pidx * kSM + trow] = st.m;
            named_bar_sync(qbar, 32 * SP);
            float mall = st.m;
#pragma unroll
            for (int pt = 0; pt < SP; ++pt) mall = fminf(mall, mfin[(pbase + pt) * kSM + trow]);
            const float thr_fin = mall + brow;
            {
                int kept = 0;
                for (int i = 0; i < st.cnt; ++i) {
                    const float v = ls[i * kSM];
                    if (v <= thr_fin) { lc[kept * kSM] = lc[i * kSM]; ++kept; }
                }
                // -1: a score inside the final threshold was LOST (or the scores were non-finite): exhaustive scan;
                // 0x100: the spill buffer holds candidates of this latent
                int nres = (sd[kSM] > thr_fin) ? (kept | (sd[0] <= thr_fin ? 0x100 : 0)) : -1;
                if (!(thr_fin < INF)) {
                    // non-finite latent (NaN / Inf in x, or scores that overflowed): every distance is NaN or
                    // infinite and the reference's argmax degenerates; take code 0 instead of an exhaustive scan
                    lc[0] = 0;
                    nres = half == 0 ? 1 : 0;
                }
                ncnt[pidx * kSM + trow] = nres;
            }
            named_bar_sync(qbar, 32 * SP);
            if (half == 0) thrfin[trow] = thr_fin;            // (every part has read mfin)
            if (NV == 1 && half == 0) {
                // merged record of this thread's latent, built here (one thread per latent, in parallel) instead of by the
                // apply warps: the count (1..4; 0 = general path) and the candidates in part order
                int tot = 0, bad = 0, cn[SP];
#pragma unroll
                for (int pt = 0; pt < SP; ++pt) {
                    const int nr = ncnt[(pbase + pt) * kSM + trow];
                    bad |= (nr < 0) | (nr & 0x100);
                    cn[pt] = nr & 0xff;
                    tot += cn[pt];
                }
                const int ncr = (bad || tot == 0 || tot > 4) ? 0 : tot;
                int cv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int jj = j < ncr ? j : 0, part_j = 0;
#pragma unroll
                    for (int pt = 0; pt < SP - 1; ++pt)
                        if (part_j == pt && jj >= cn[pt]) { jj -= cn[pt]; part_j = pt + 1; }
                    const int c = cand_c[((pbase + part_j) * kSCand + (jj < kSCand ? jj : 0)) * kSM + trow];
                    cv[j] = (c >= 0 && c < p.k) ? c : 0;
                }
                recc[trow] = make_int4(cv[0], cv[1], cv[2], cv[3]);
                recn[trow] = ncr;
            }
            named_bar_sync(qbar, 32 * SP);
            SP_LAP(4);
            // ---- resolve + apply: all 32 lanes on one latent, R latents in flight, ONE round trip to L2:
            //      x rows, the first candidate's code word and (for ambiguous latents) candidates 2-4 are
            //      all requested before anything is used.  This warp takes 16 latents of its quadrant.
            const int lrow0 = quad * 32 + half * LPW;         // first latent (within the tile) of this warp
            const int64_t wrow0 = (int64_t)tile * kSM + lrow0;
            int mycode = 0;
            float loss = 0.f;
            unsigned gen_mask = 0;                            // latents left to the general path (second pass)
            if constexpr (NV == 1) {
                // d <= 128.  Lane i looks up latent i of this warp once; three passes:
                //   1. one candidate (65 - 80 % of the latents): nothing to decide — x row and code word of EVERY such latent
                //      are requested before the first is used (one L2 round trip for the whole warp); at d <= 64 a row is 16
                //      lanes wide, so one warp-wide access serves two latents;
                //   2. two to four candidates: fp32 re-score (fp64 if inseparable), two latents per round trip;
                //   3. anything else (second pass below).
                int nci = -1;
                if (lane < LPW && wrow0 + lane < p.n) { nci = recn[lrow0 + lane]; mycode = recc[lrow0 + lane].x; }
                const unsigned m1 = __ballot_sync(0xffffffffu, nci == 1);
                unsigned m2 = __ballot_sync(0xffffffffu, nci >= 2);
                gen_mask = __ballot_sync(0xffffffffu, nci == 0);
                constexpr int LPI = DP == 64 ? 2 : 1;         // latents per warp-wide 16-byte access
                constexpr int NACC = LPW / LPI;               // accesses that cover the warp's latents
                constexpr int GU = NACC < 8 ? NACC : 8;       // ... in flight at a time
                const int hs = LPI == 2 ? (lane >> 4) : 0;
                const int c4 = LPI == 2 ? (lane & 15) : lane;
                const bool hv = c4 < nchunk;
                const float4* cb4 = reinterpret_cast<const float4*>(p.cb);
                const int dq = p.d >> 2;
#pragma unroll 1
                for (int g0 = 0; g0 < NACC; g0 += GU) {
                    if (!TRAIN && p.q == nullptr) break;      // assignment only: a one-candidate latent needs nothing more
                    if (((m1 >> (g0 * LPI)) & ((1u << (GU * LPI)) - 1u)) == 0u) continue;     // warp-uniform
                    float4 xa[GU], ea[GU];
                    int cu[GU];
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        const int lat = (g0 + u) * LPI + hs;
                        xa[u] = z4; ea[u] = z4; cu[u] = 0;
                        if ((m1 >> lat) & 1u) {
                            cu[u] = recc[lrow0 + lat].x;
                            if (hv) {
                                if (TRAIN) xa[u] = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)(wrow0 + lat) * p.d) + c4);   // (eval: q = e)
                                ea[u] = __ldg(cb4 + (size_t)cu[u] * dq + c4);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        const int lat = (g0 + u) * LPI + hs;
                        if (((m1 >> lat) & 1u) && (p.q != nullptr || TRAIN))
                            apply_row<1, TRAIN>(p, esum, xa[u], z4, ea[u], z4, cu[u], wrow0 + lat, hv, false, c4, loss);
                    }
                }
#pragma unroll 1
                while (m2) {                                   // warp-uniform
                    constexpr int R2 = 2;
                    float4 xa[R2], ea[R2][4];
                    float e2v[R2][4];
                    int nc[R2], cc[R2][4], lat[R2];
#pragma unroll
                    for (int u = 0; u < R2; ++u) {
                        nc[u] = 0; lat[u] = 0;
                        if (m2) { lat[u] = __ffs((int)m2) - 1; m2 &= m2 - 1u; nc[u] = recn[lrow0 + lat[u]]; }
                        const int4 c4v = recc[lrow0 + lat[u]];
                        cc[u][0] = c4v.x; cc[u][1] = c4v.y; cc[u][2] = c4v.z; cc[u][3] = c4v.w;
                        const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)(wrow0 + lat[u]) * p.d);
                        xa[u] = (nc[u] > 0 && h0) ? __ldg(xr + lane) : z4;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ea[u][j] = z4; e2v[u][j] = 0.f;
                            if (j < nc[u]) {
                                if (h0) ea[u][j] = __ldg(cb4 + (size_t)cc[u][j] * dq + lane);
                                e2v[u][j] = __ldg(p.e2 + cc[u][j]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < R2; ++u) {
                        if (nc[u] < 2) continue;               // (second slot of an odd batch)
                        // fp32 re-score of the (<= 4) candidates, all lanes on this latent
                        n_resc += 1u;
                        int sel = 0;
                        float dd[4];
                        float ss = fmaf(xa[u].x, xa[u].x, fmaf(xa[u].y, xa[u].y, fmaf(xa[u].z, xa[u].z, xa[u].w * xa[u].w)));
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dd[j] = fmaf(xa[u].x, ea[u][j].x, fmaf(xa[u].y, ea[u][j].y, fmaf(xa[u].z, ea[u][j].z, xa[u].w * ea[u][j].w)));
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            ss += __shfl_xor_sync(0xffffffffu, ss, off);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], off);
                        }
                        const float bnd = fmaf(sqrtf(ss), 1.0001f, emax);
                        const float thr32 = 1.6e-6f * bnd * bnd;
                        float mm1 = INF, mm2 = INF;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float sv = j < nc[u] ? fmaf(-2.f, dd[j], e2v[u][j]) : INF;
                            if (sv < mm1) { mm2 = mm1; mm1 = sv; sel = j; }
                            else if (sv < mm2) mm2 = sv;
                        }
                        if (!(mm2 - mm1 > thr32)) {
                            // canonical fp64 rule among the candidates (code words already in registers)
                            n_f64 += 1u;
                            double pp = 0.0;
                            if (h0) pp = dot4(pp, xa[u], xa[u]);
                            const float x2 = __double2float_rn(butterfly_sum(pp));
                            float best = INF;
                            int arg = 0x7fffffff;
                            sel = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (j < nc[u]) {
                                    double sdd = 0.0;
                                    if (h0) sdd = dot4(sdd, xa[u], ea[u][j]);
                                    const float dk = canon_score(x2, butterfly_sum(sdd), e2v[u][j]);
                                    if (dk < best || (dk == best && cc[u][j] < arg)) { best = dk; arg = cc[u][j]; sel = j; }
                                }
                            }
                        }
                        const int code = sel == 0 ? cc[u][0] : sel == 1 ? cc[u][1] : sel == 2 ? cc[u][2] : cc[u][3];
                        const float4 wa = sel == 0 ? ea[u][0] : sel == 1 ? ea[u][1] : sel == 2 ? ea[u][2] : ea[u][3];
                        if (p.q != nullptr || TRAIN)
                            apply_row<1, TRAIN>(p, esum, xa[u], z4, wa, z4, code, wrow0 + lat[u], h0, false, lane, loss);
                        mycode = (lane == lat[u]) ? code : mycode;
                    }
                }
            } else {
                // d = 256: the same three passes, but the merged record of latent i lives in the registers of lane i (its
                // 5 KB in shared memory would cost the fourth code-slab stage) and is handed out by shuffles.
                int rn = -1;
                int rc0 = 0, rc1 = 0, rc2 = 0, rc3 = 0;
                if (lane < LPW && wrow0 + lane < p.n) {
                    const int lrow = lrow0 + lane;
                    int tot = 0, bad = 0, cn[SP];
#pragma unroll
                    for (int pt = 0; pt < SP; ++pt) {
                        const int nr = ncnt[(pbase + pt) * kSM + lrow];
                        bad |= (nr < 0) | (nr & 0x100);
                        cn[pt] = nr & 0xff;
                        tot += cn[pt];
                    }
                    rn = (bad || tot == 0 || tot > 4) ? 0 : tot;
                    int cv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int jj = j < rn ? j : 0, part_j = 0;
#pragma unroll
                        for (int pt = 0; pt < SP - 1; ++pt)
                            if (part_j == pt && jj >= cn[pt]) { jj -= cn[pt]; part_j = pt + 1; }
                        const int c = cand_c[((pbase + part_j) * kSCand + (jj < kSCand ? jj : 0)) * kSM + lrow];
                        cv[j] = (c >= 0 && c < p.k) ? c : 0;
                    }
                    rc0 = cv[0]; rc1 = cv[1]; rc2 = cv[2]; rc3 = cv[3];
                    mycode = rc0;
                }
                const unsigned m1 = __ballot_sync(0xffffffffu, rn == 1);
                unsigned m2 = __ballot_sync(0xffffffffu, rn >= 2);
                gen_mask = __ballot_sync(0xffffffffu, rn == 0);
                const float4* cb4 = reinterpret_cast<const float4*>(p.cb);
                const int dq = p.d >> 2;
                constexpr int GU = 4;                         // one-candidate latents in flight (16 registers each)
#pragma unroll 1
                for (int g0 = 0; g0 < LPW; g0 += GU) {
                    if (!TRAIN && p.q == nullptr) break;      // assignment only: a one-candidate latent needs nothing more
                    if (((m1 >> g0) & ((1u << GU) - 1u)) == 0u) continue;     // warp-uniform
                    float4 xa[GU], xb[GU], ea[GU], eb[GU];
                    int cu[GU];
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        const int lat = g0 + u;
                        cu[u] = __shfl_sync(0xffffffffu, rc0, lat);
                        xa[u] = z4; xb[u] = z4; ea[u] = z4; eb[u] = z4;
                        if ((m1 >> lat) & 1u) {
                            const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)(wrow0 + lat) * p.d);
                            const float4* er = cb4 + (size_t)cu[u] * dq;
                            if (h0) { if (TRAIN) xa[u] = __ldg(xr + lane); ea[u] = __ldg(er + lane); }
                            if (h1) { if (TRAIN) xb[u] = __ldg(xr + lane + 32); eb[u] = __ldg(er + lane + 32); }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        const int lat = g0 + u;
                        if ((m1 >> lat) & 1u)
                            apply_row<NV, TRAIN>(p, esum, xa[u], xb[u], ea[u], eb[u], cu[u], wrow0 + lat, h0, h1, lane, loss);
                    }
                }
#pragma unroll 1
                while (m2) {                                   // warp-uniform
                    constexpr int R2 = 2;
                    float4 xa[R2], xb[R2], ea[R2][4], eb[R2][4];
                    float e2v[R2][4];
                    int nc[R2], cc[R2][4], lat[R2];
#pragma unroll
                    for (int u = 0; u < R2; ++u) {
                        nc[u] = 0; lat[u] = 0;
                        if (m2) { lat[u] = __ffs((int)m2) - 1; m2 &= m2 - 1u; nc[u] = 1; }
                        const int nrec = __shfl_sync(0xffffffffu, rn, lat[u]);
                        nc[u] = nc[u] ? nrec : 0;
                        cc[u][0] = __shfl_sync(0xffffffffu, rc0, lat[u]); cc[u][1] = __shfl_sync(0xffffffffu, rc1, lat[u]);
                        cc[u][2] = __shfl_sync(0xffffffffu, rc2, lat[u]); cc[u][3] = __shfl_sync(0xffffffffu, rc3, lat[u]);
                        const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)(wrow0 + lat[u]) * p.d);
                        xa[u] = (nc[u] > 0 && h0) ? __ldg(xr + lane) : z4;
                        xb[u] = (nc[u] > 0 && h1) ? __ldg(xr + lane + 32) : z4;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ea[u][j] = z4; eb[u][j] = z4; e2v[u][j] = 0.f;
                            if (j < nc[u]) {
                                const float4* er = cb4 + (size_t)cc[u][j] * dq;
                                if (h0) ea[u][j] = __ldg(er + lane);
                                if (h1) eb[u][j] = __ldg(er + lane + 32);
                                e2v[u][j] = __ldg(p.e2 + cc[u][j]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < R2; ++u) {
                        if (nc[u] < 2) continue;               // (second slot of an odd batch)
                        n_resc += 1u;
                        int sel = 0;
                        float dd[4];
                        float ss = fmaf(xa[u].x, xa[u].x, fmaf(xa[u].y, xa[u].y, fmaf(xa[u].z, xa[u].z, xa[u].w * xa[u].w)));
                        ss = fmaf(xb[u].x, xb[u].x, fmaf(xb[u].y, xb[u].y, fmaf(xb[u].z, xb[u].z, fmaf(xb[u].w, xb[u].w, ss))));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dd[j] = fmaf(xa[u].x, ea[u][j].x, fmaf(xa[u].y, ea[u][j].y, fmaf(xa[u].z, ea[u][j].z, xa[u].w * ea[u][j].w)));
                            dd[j] = fmaf(xb[u].x, eb[u][j].x, fmaf(xb[u].y, eb[u][j].y, fmaf(xb[u].z, eb[u][j].z, fmaf(xb[u].w, eb[u][j].w, dd[j]))));
                        }
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            ss += __shfl_xor_sync(0xffffffffu, ss, off);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], off);
                        }
                        const float bnd = fmaf(sqrtf(ss), 1.0001f, emax);
                        const float thr32 = 1.6e-6f * bnd * bnd;
                        float mm1 = INF, mm2 = INF;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float sv = j < nc[u] ? fmaf(-2.f, dd[j], e2v[u][j]) : INF;
                            if (sv < mm1) { mm2 = mm1; mm1 = sv; sel = j; }
                            else if (sv < mm2) mm2 = sv;
                        }
                        if (!(mm2 - mm1 > thr32)) {
                            // canonical fp64 rule among the candidates (code words already in registers)
                            n_f64 += 1u;
                            double pp = 0.0;
                            if (h0) pp = dot4(pp, xa[u], xa[u]);
                            if (h1) pp = dot4(pp, xb[u], xb[u]);
                            const float x2 = __double2float_rn(butterfly_sum(pp));
                            float best = INF;
                            int arg = 0x7fffffff;
                            sel = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (j < nc[u]) {
                                    double sdd = 0.0;
                                    if (h0) sdd = dot4(sdd, xa[u], ea[u][j]);
                                    if (h1) sdd = dot4(sdd, xb[u], eb[u][j]);
                                    const float dk = canon_score(x2, butterfly_sum(sdd), e2v[u][j]);
                                    if (dk < best || (dk == best && cc[u][j] < arg)) { best = dk; arg = cc[u][j]; sel = j; }
                                }
                            }
                        }
                        const int code = sel == 0 ? cc[u][0] : sel == 1 ? cc[u][1] : sel == 2 ? cc[u][2] : cc[u][3];
                        const float4 wa = sel == 0 ? ea[u][0] : sel == 1 ? ea[u][1] : sel == 2 ? ea[u][2] : ea[u][3];
                        const float4 wb = sel == 0 ? eb[u][0] : sel == 1 ? eb[u][1] : sel == 2 ? eb[u][2] : eb[u][3];
                        if (p.q != nullptr || TRAIN)
                            apply_row<NV, TRAIN>(p, esum, xa[u], xb[u], wa, wb, code, wrow0 + lat[u], h0, h1, lane, loss);
                        mycode = (lane == lat[u]) ? code : mycode;
                    }
                }
            }
            // ---- second pass (rare): long merged lists, spilled candidates, exhaustive scans — one latent at a time
            SP_MARK(sp_g0);
#pragma unroll 1
            while (gen_mask) {
                const int r = __ffs(gen_mask) - 1;
                gen_mask &= gen_mask - 1;
                const int lrow = lrow0 + r;
                const int64_t grow = wrow0 + r;
                const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)grow * p.d);
                const float4 xa = h0 ? __ldg(xr + lane) : z4;
                const float4 xb = h1 ? __ldg(xr + lane + 32) : z4;
                int cn[SP], badg = 0;
#pragma unroll
                for (int pt = 0; pt < SP; ++pt) {
                    const int nr = ncnt[(pbase + pt) * kSM + lrow];
                    badg |= nr < 0;
                    cn[pt] = nr & 0xff;
                }
                int ncg = -1;
                if (!badg) {
                    ncg = gather_cands<SP>(scratch, cn, cand_c + pbase * (kSCand * kSM) + lrow, ob, lrow, thrfin[lrow], lane);
                    if (ncg == 0) ncg = -1;
                }
                const int rr = resolve_stream<NV>(xa, xb, ncg, scratch, p.cb, p.e2, p.k, p.d, emax, lane);
                __syncwarp();
                n_resc += 1u;
                n_f64 += (unsigned)(rr >> 30);
                int code = rr & 0x3fffffff;
                code = code < p.k ? code : 0;
                if (p.q != nullptr || TRAIN) {
                    const float4* er = reinterpret_cast<const float4*>(p.cb + (size_t)code * p.d);
                    const float4 wa = h0 ? __ldg(er + lane) : z4;
                    const float4 wb = h1 ? __ldg(er + lane + 32) : z4;
                    apply_row<NV, TRAIN>(p, esum, xa, xb, wa, wb, code, grow, h0, h1, lane, loss);
                }
                mycode = (lane == r) ? code : mycode;
            }
            SP_ADD(7, sp_g0);                                  // second pass (general resolution, one latent at a time)
            if (lane < LPW && wrow0 + lane < p.n) {
                p.idx[wrow0 + lane] = (int64_t)mycode;
                atomicAdd(p.stats + mycode, 1.0f);            // counts (exact integers in fp32)
            }
            loss_d += (double)loss;
            named_bar_sync(qbar, 32 * SP);                // the quadrant's lists are reused by the next row tile
            SP_LAP(5);
        }
        if (warp == 4 && lane == 0) SP_DUMP(16);
        if (lane == 0) SP_DUMP(64 + 8 * (warp - 4));
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();                      // the peer may read this CTA's shared memory until its last MMA is done
    tc_fence_after();
    if (warp == 1) { if (CG == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
    if (lane == 0 && n_resc) {
        atomicAdd(&p.hdr->n_rescored, n_resc);
        if (n_f64) atomicAdd(&p.hdr->n_exact, n_f64);
    }
    if (TRAIN) {
        const double t = block_sum(loss_d, red);
        if (tid == 0) atomicAdd(&p.hdr->loss_sum, t);
    }
    finish_ticket<TRAIN>(p, red, misc);
}

}  // namespace tvq
