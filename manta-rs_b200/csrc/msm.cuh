// Internal interface of the bucket-method MSM (msm.cu) used by the stand-alone entry points and by the prover.
//
// Replaces ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul`.  Pipeline per scalar list:
//   digits+histogram -> scan/plan -> scatter (counting sort by bucket) -> bucket accumulate
//   -> 2-level bucket reduction (running sums, then a row/column split of the chunk index) -> (optional) Horner.
// Bucket accumulation has two implementations:
//   * batched affine (default): every bucket is summed by a pairwise tree, one round per tree level; all pairs of a
//     round (every bucket, MSM and proof of the launch) are independent affine additions that share their field
//     inversion by Montgomery's trick (thread-local prefix products -> 8..64-way second level -> one word-level binary-Euclid
//     inversion per 64 threads on the ALU pipe).  6 field multiplications per addition instead of 10.
//   * XYZZ (MP_MSM_XYZZ=1): one thread per bucket slice, mixed additions into an extended-Jacobian accumulator.
// Signed c-bit digits: 2^(c-1) buckets per window group.  A "table" holds `rows` precomputed multiples
// 2^(c*groups*t) * P_i (t < rows) so that window w = t*groups + g lands in bucket set g with point row t;
// rows == windows (groups == 1) removes the Horner step entirely (used for the circuit keys).
#pragma once
#include "common.cuh"

namespace mp {

constexpr int MSM_CLASSES = 64;        // length classes used to order work items (longest first)
constexpr int MSM_MAX_JOBS = 6;        // MSMs handled by one accumulate / reduce launch
constexpr int MSM_HEAVY_SEGS = 33;     // buckets with this many slices or more are folded by a whole warp
constexpr int PLAN_THREADS = 1024;     // threads of the per-list plan block = bucket chunks of the pair index
constexpr int BA_BLK = 128;            // threads per block of the batched-affine round kernels
constexpr int BA_T2_MAX = 64;          // most thread totals one second-level inversion thread takes (chosen per round)
constexpr int BA_MAX_ROUNDS = 28;

bool msm_use_batched_affine();         // false when MP_MSM_XYZZ=1 is set in the environment

struct MsmGeom {
    int c;            // window bits (2..16)
    int windows;      // 255 / c + 1
    int groups;       // bucket sets
    int rows;         // table rows = ceil(windows / groups)
    uint32_t n_scalars;      // scalars per list
    uint32_t table_stride;   // points per table row (>= n_scalars)
    uint32_t bpg;            // buckets per group = 2^(c-1)
    uint32_t n_buckets;      // groups * bpg
    uint32_t max_entries;    // n_scalars * windows
    uint32_t seg;            // max entries one thread accumulates (power of two >= 64, ~2x the mean bucket load)
    uint32_t max_items;      // n_buckets + max_entries / seg + 1
    uint32_t max_heavy;      // max_entries / (seg * (MSM_HEAVY_SEGS - 1)) + 1
    uint32_t red_s1;         // buckets per thread in reduction level 1
    uint32_t l1pg;           // level-1 chunks per group = ceil(bpg / red_s1)
    uint32_t red_d;          // level 2 works on a red_d x red_d grid of level-1 chunks (power of two, red_d^2 >= l1pg)
    // batched-affine accumulation
    uint32_t ent_cap;        // entry slots per list: max_entries + n_buckets (bucket starts are padded to even)
    uint32_t p_cap;          // affine point slots per list = ent_cap / 2 (bucket k owns slots from start[k] / 2)
    uint32_t plan_per;       // buckets per plan chunk = ceil(n_buckets / PLAN_THREADS)
    int ba_rounds;           // tree levels needed for the fullest possible bucket
};
// batch_hint: expected number of independent scalar vectors per launch (sizes the reduction chunks)
MsmGeom msm_geom(int c, int groups, uint32_t n_scalars, uint32_t table_stride, size_t batch_hint);

// Per-list sort workspace for a batch of `batch` independent scalar vectors.
struct MsmSortWs {
    uint32_t* cnt = nullptr;        // [batch][n_buckets]   entries per bucket
    uint32_t* start = nullptr;      // [batch][n_buckets]   first entry of bucket
    uint32_t* slot_base = nullptr;  // [batch][n_buckets+1] first partial-sum slot of bucket
    uint32_t* fill = nullptr;       // [batch][n_buckets]   scatter cursors
    uint32_t* items = nullptr;      // [batch][max_items]   work items (bucket, segment), longest first
    uint32_t* n_items = nullptr;    // [batch]
    uint32_t* entries = nullptr;    // [batch][max_entries] sorted (sign << 31 | table index)
    uint32_t* heavy = nullptr;      // [batch][max_heavy]   buckets split into >= MSM_HEAVY_SEGS slices
    uint32_t* n_heavy = nullptr;    // [batch]
    uint32_t* q = nullptr;          // [batch][ba_rounds+1][PLAN_THREADS+1] pairs of round r before chunk (batched affine)
    size_t bytes(const MsmGeom& g, size_t batch) const;
};
int msm_sort_ws_alloc(MsmSortWs& ws, const MsmGeom& g, size_t batch, DevBuf& backing);

// scalars: [batch][scalar_stride] canonical Fr (8 x u32 each); only the first n_scalars of each row are used.
// valid (nullable): bitmap over scalar indices; scalars whose base is the point at infinity are dropped here so
// that they never occupy a lane of the accumulation kernel.
int msm_sort(const MsmGeom& g, const uint32_t* scalars, size_t scalar_stride_words, size_t batch, const MsmSortWs& ws,
             const uint32_t* valid, cudaStream_t st);

// One MSM of a launch: a table, the sorted list it consumes, and its outputs.
struct MsmJob {
    MsmGeom g;
    MsmSortWs ws;
    const void* table;   // Affine<F>[rows][table_stride]
    void* partial;       // [batch][max_items] XYZZ<F> partial sums, indexed by slot
    void* result;        // [batch][groups] XYZZ<F> group results after reduction
    void* scratch;       // reduction scratch, msm_reduce_scratch_bytes()
    void* pbuf;          // [batch][p_cap] Affine<F>: tree levels of the batched-affine accumulation, bucket sums at the end
    // batched-affine bucket reduction (row / column sums of the bucket grid, see msm_geom_rc)
    MsmGeom g_rc;        // geometry of the row/column stage
    MsmSortWs ws_rc;     // its segment lists (cnt, start, entries, q)
    void* pbuf_rc;       // [batch][g_rc.p_cap] Affine<F>
    void* result_rc;     // [batch][g_rc.groups] XYZZ<F>: weighted column / row sums per window group
};
// Row/column stage of the reduction  sum_k (k+1) B_k  over the 2^(c-1) buckets of a window group: with k+1 = 256 hi + lo,
//   sum = 256 * sum_hi hi * R_hi + sum_lo lo * C_lo,   R_hi / C_lo = sums of the buckets in row hi / column lo.
// The 2 x 2^(c-1) additions of the row and column sums are pairwise trees (batched affine, same kernels as the bucket
// accumulation); the two weighted sums over <= 256 points run through the running-sum kernels on a 2-group geometry.
// Segment s of group g: [g*512, g*512+255) = columns lo = 1..255, [g*512+256, g*512+512) = rows hi = 1..256.
MsmGeom msm_geom_rc(const MsmGeom& g);
int msm_rc_plan(const MsmGeom& g, const MsmSortWs& ws, const MsmGeom& g_rc, const MsmSortWs& ws_rc, size_t batch, cudaStream_t st);
// Scratch of one batched-affine launch (all jobs x batch of the launch share one inversion tree per round).
struct MsmBaWs {
    void* prefix = nullptr;    // Fq per pair: thread-local prefix products of the denominators (G2: of their norms)
    uint32_t* desc = nullptr;  // 3 x u32 per pair: sources and destination
    void* tot = nullptr;       // Fq per thread: product of its denominators, then its inverse
    void* pre2 = nullptr;      // Fq per thread: second-level prefix products
    void* norm = nullptr;      // G2 only: Fq per pair, the norm of the denominator (forward pass -> backward pass)
    size_t cap_pairs = 0, cap_threads = 0;  // slots behind prefix/desc and tot/pre2 (sized for every count <= capacity)
    // Software pipeline over the slabs of a tree level (batches of more than 32 vectors): a SECOND scratch set of the same size,
    // so that the forward pass of slab i+1 and the backward pass of slab i-1 run while the latency-bound inversion kernel of
    // slab i sits on a high-priority side stream.  n_sets == 1 or no side streams: one launch triple after the other.
    int n_sets = 1;
    void* prefix_b = nullptr;
    uint32_t* desc_b = nullptr;
    void* tot_b = nullptr;
    void* pre2_b = nullptr;
    void* norm_b = nullptr;
    cudaStream_t mid_st[2] = {nullptr, nullptr};                               // side streams of k_ba_mid (owned by the caller)
    cudaEvent_t ev_fwd[2] = {nullptr, nullptr}, ev_mid[2] = {nullptr, nullptr};  // fwd done -> mid may start; mid done -> bwd may start
    mutable size_t dom_count = 0;           // vectors covered by the launch the ev_bwd0/1 events bracket (first slab of level 1)
    int round_limit = 0;                    // > 0: launch only this many tree levels (msm_ba_rounds_needed), 0: every provisioned level
    // (a limit below the populated levels leaves up to 2^(populated - limit) points per bucket; the latency path of the bucket
    //  reduction, reduce_small, adds those while it walks the buckets - see msm_ba_trim_levels)
    // optional: recorded around the round-1 k_ba_bwd launch of the bucket trees (the dominant kernel of a proof)
    cudaEvent_t ev_bwd0 = nullptr, ev_bwd1 = nullptr;
};
// slab: vectors the scratch is sized for (0 = msm_ba_slab(batch); a caller with memory to spare passes more)
size_t msm_ba_ws_bytes(const MsmGeom* geoms, int n_jobs, size_t batch, bool g2, size_t slab = 0);
void msm_ba_ws_bind(MsmBaWs& ws, const MsmGeom* geoms, int n_jobs, size_t batch, bool g2, void* mem, size_t slab = 0);
size_t msm_ba_slab(size_t batch, bool g2);   // vectors the round scratch of a batch of `batch` is sized for
// host-only: {pairs, threads} a live `count` needs and {pairs, threads} a workspace sized for `cap` provides
void msm_ba_ws_demand(const MsmGeom* geoms, int n_jobs, size_t cap, size_t count, uint64_t out[4]);
int msm_ba_rounds_needed(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st, int* out_rounds);
// Tree levels a latency-bound launch (<= 2 vectors) leaves to the reduction (MP_BA_TRIM_LEVELS, default 0): every level costs a
// serial ~85 us inversion, the <= 2^trim leftover points of a bucket cost the reduction's lanes serial mixed additions.
// Measured on the B200 (profiles/r02q_*): 4.89 / 5.23 / 5.62 / 7.64 ms per proof for 0 / 1 / 2 / 3 levels - the additions cost
// more than the levels they replace - so nothing is trimmed by default.
int msm_ba_trim_levels();
// ba == nullptr selects the XYZZ accumulation (jobs[].partial), otherwise batched affine (jobs[].pbuf)
int msm_accumulate_g1(const MsmJob* jobs, int n_jobs, size_t batch, const MsmBaWs* ba, cudaStream_t st);
int msm_accumulate_g2(const MsmJob* jobs, int n_jobs, size_t batch, const MsmBaWs* ba, cudaStream_t st);
size_t msm_reduce_scratch_bytes(const MsmGeom& g, size_t batch, bool g2);
// The reduction in two parts, so that the caller can put the latency-bound tail on another stream:
//   heavy: throughput work (XYZZ: level-1 running sums; batched affine: row/column trees)
//   tail : the small weighted sums (level 2 / the 2-group running sums) -> jobs[].result
int msm_reduce_heavy_g1(const MsmJob* jobs, int n_jobs, size_t batch, const MsmBaWs* ba, cudaStream_t st);
int msm_reduce_heavy_g2(const MsmJob* jobs, int n_jobs, size_t batch, const MsmBaWs* ba, cudaStream_t st);
int msm_reduce_tail_g1(const MsmJob* jobs, int n_jobs, size_t batch, const MsmBaWs* ba, cudaStream_t st);
int msm_reduce_tail_g2(const MsmJob* jobs, int n_jobs, size_t batch, const MsmBaWs* ba, cudaStream_t st);

// bitmap[i] = base i is not the point at infinity (optionally OR-ed into an existing bitmap)
int msm_validity_g1(const void* bases, uint32_t n, uint32_t* bitmap, bool accumulate, cudaStream_t st);
int msm_validity_g2(const void* bases, uint32_t n, uint32_t* bitmap, bool accumulate, cudaStream_t st);

// Table construction: rows multiples of every base, 2^(c*groups*t) * P_i, affine, Montgomery.
// bases: Affine<F>[n] on the device (already Montgomery); out: Affine<F>[rows][stride]; entries >= n are infinity.
int msm_build_table_g1(const MsmGeom& g, const void* bases, uint32_t n, void* out, cudaStream_t st);
int msm_build_table_g2(const MsmGeom& g, const void* bases, uint32_t n, void* out, cudaStream_t st);

// out[i] = scalars[i] * G (standard generator), device-resident: canonical Fr in, Affine<F> Montgomery out
int msm_fixed_base_dev_g1(const void* d_scalars, size_t n, void* d_out, cudaStream_t st);
int msm_fixed_base_dev_g2(const void* d_scalars, size_t n, void* d_out, cudaStream_t st);

// Horner over window groups: out = sum_g 2^(c*g) * result[g]  (one thread per MSM; only when groups > 1)
int msm_horner_g1(const MsmGeom& g, const void* group_results, void* out_xyzz, size_t count, cudaStream_t st);
int msm_horner_g2(const MsmGeom& g, const void* group_results, void* out_xyzz, size_t count, cudaStream_t st);

}  // namespace mp
