// Bucket-method multi-scalar multiplication on sm_100a (see msm.cuh for the pipeline).
//
// Replaces ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` — the five calls inside ark-groth16's
// `create_proof` behind manta-crypto/src/arkworks/groth16.rs:597 and the direct benchmark use at
// manta-benchmark/src/ecc.rs:62-118.  Only the affine value of the result is observable,
// so signed digits, precomputed window tables and XYZZ buckets give bit-identical outputs.
#include <algorithm>
#include <vector>

#include "msm.cuh"

namespace mp {

// ---------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------
MsmGeom msm_geom(int c, int groups, uint32_t n_scalars, uint32_t table_stride, size_t batch_hint) {
    MsmGeom g{};
    g.c = c;
    g.windows = 255 / c + 1;
    if (groups <= 0 || groups > g.windows) groups = g.windows;
    g.groups = groups;
    g.rows = (g.windows + groups - 1) / groups;
    g.n_scalars = n_scalars;
    g.table_stride = table_stride;
    g.bpg = 1u << (c - 1);
    g.n_buckets = g.bpg * groups;
    g.max_entries = n_scalars * (uint32_t)g.windows;
    // slice length: ~2x the mean bucket load, so that uniform inputs give one slice per bucket and only skewed
    // buckets are split
    g.seg = 64;
    while (g.seg < 4096 && (uint64_t)g.seg * g.n_buckets < 2ull * g.max_entries) g.seg *= 2;
    g.max_items = g.n_buckets + g.max_entries / g.seg + 1;
    g.max_heavy = g.max_entries / (g.seg * (MSM_HEAVY_SEGS - 1)) + 1;
    // Reduction level 1 is throughput work (2 adds per bucket); its chunk length only sets how many threads share
    // it.  Aim at >= ~64k threads per launch (148 SMs x a few hundred resident threads); `batch_hint` counts the
    // (job x vector) instances that share the launch.  Longer chunks shorten the latency-bound level 2.
    size_t inst = std::max<size_t>(batch_hint, 1) * (size_t)groups;
    size_t want = 65536;
    size_t s1 = (size_t)g.bpg * inst / want;
    g.red_s1 = 8;
    while (g.red_s1 * 2 <= s1 && g.red_s1 < 128) g.red_s1 *= 2;
    if (g.red_s1 > g.bpg) g.red_s1 = g.bpg;
    g.l1pg = (g.bpg + g.red_s1 - 1) / g.red_s1;
    g.red_d = 1;
    while (g.red_d * g.red_d < g.l1pg) g.red_d <<= 1;
    return g;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

size_t MsmSortWs::bytes(const MsmGeom& g, size_t batch) const {
    size_t b = 0;
    b += align256(batch * g.n_buckets * 4) * 3;          // cnt, start, fill
    b += align256(batch * (g.n_buckets + 1) * 4);        // slot_base
    b += align256(batch * (size_t)g.max_items * 8);      // items
    b += align256(batch * 4);                            // n_items
    b += align256(batch * (size_t)g.max_entries * 4);    // entries
    b += align256(batch * (size_t)g.max_heavy * 4);      // heavy
    b += align256(batch * 4);                            // n_heavy
    return b;
}

int msm_sort_ws_alloc(MsmSortWs& ws, const MsmGeom& g, size_t batch, DevBuf& backing) {
    MP_TRY(backing.alloc(ws.bytes(g, batch)));
    char* p = backing.as<char>();
    auto take = [&](size_t n) { char* r = p; p += align256(n); return r; };
    ws.cnt = (uint32_t*)take(batch * g.n_buckets * 4);
    ws.start = (uint32_t*)take(batch * g.n_buckets * 4);
    ws.fill = (uint32_t*)take(batch * g.n_buckets * 4);
    ws.slot_base = (uint32_t*)take(batch * (g.n_buckets + 1) * 4);
    ws.items = (uint32_t*)take(batch * (size_t)g.max_items * 8);
    ws.n_items = (uint32_t*)take(batch * 4);
    ws.entries = (uint32_t*)take(batch * (size_t)g.max_entries * 4);
    ws.heavy = (uint32_t*)take(batch * (size_t)g.max_heavy * 4);
    ws.n_heavy = (uint32_t*)take(batch * 4);
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// signed-digit recoding
// ---------------------------------------------------------------------------------------------------------
// bits [pos, pos + c) of a 256-bit little-endian integer (c <= 16)
MP_DEV uint32_t scalar_bits(const uint32_t* s, int pos, int c) {
    int w = pos >> 5, o = pos & 31;
    uint64_t v = s[w];
    if (w + 1 < 8) v |= (uint64_t)s[w + 1] << 32;
    return (uint32_t)(v >> o) & ((1u << c) - 1);
}

// Calls f(window, bucket_in_group (0-based), negative) for every non-zero digit.
template <class Fn>
MP_DEV void for_each_digit(const uint32_t* s, int c, int windows, Fn f) {
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (int w = 0; w < windows; w++) {
        uint32_t raw = scalar_bits(s, w * c, c) + carry;
        bool neg = raw > half;
        uint32_t mag = neg ? (1u << c) - raw : raw;
        carry = neg ? 1u : 0u;
        if (mag) f(w, mag - 1, neg);
    }
}

__global__ void k_msm_hist(MsmGeom g, const uint32_t* __restrict__ scalars, size_t stride_words, const uint32_t* __restrict__ valid, uint32_t* cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = blockIdx.y;
    if (i >= g.n_scalars) return;
    if (valid && !((valid[i >> 5] >> (i & 31)) & 1)) return;
    uint32_t s[8];
    const uint4* p = reinterpret_cast<const uint4*>(scalars + b * stride_words + (size_t)i * 8);
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    uint32_t* my = cnt + (size_t)b * g.n_buckets;
    for_each_digit(s, g.c, g.windows, [&](int w, uint32_t bk, bool) {
        uint32_t grp = w % g.groups;
        atomicAdd(&my[grp * g.bpg + bk], 1u);
    });
}

__global__ void k_msm_scatter(MsmGeom g, const uint32_t* __restrict__ scalars, size_t stride_words, const uint32_t* __restrict__ valid,
                              const uint32_t* __restrict__ start, uint32_t* fill, uint32_t* entries) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t b = blockIdx.y;
    if (i >= g.n_scalars) return;
    if (valid && !((valid[i >> 5] >> (i & 31)) & 1)) return;
    uint32_t s[8];
    const uint4* p = reinterpret_cast<const uint4*>(scalars + b * stride_words + (size_t)i * 8);
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    const uint32_t* st = start + (size_t)b * g.n_buckets;
    uint32_t* fl = fill + (size_t)b * g.n_buckets;
    uint32_t* en = entries + (size_t)b * g.max_entries;
    for_each_digit(s, g.c, g.windows, [&](int w, uint32_t bk, bool neg) {
        uint32_t grp = w % g.groups, row = w / g.groups;
        uint32_t k = grp * g.bpg + bk;
        uint32_t pos = st[k] + atomicAdd(&fl[k], 1u);
        en[pos] = (neg ? 0x80000000u : 0u) | (row * g.table_stride + i);
    });
}

// One block per batch element: prefix sums over bucket counts (entry offsets and partial-sum slots) and the
// work-item list ordered by descending segment length, so that the lanes of a warp run equally long loops.
constexpr int PLAN_THREADS = 1024;
__global__ void __launch_bounds__(PLAN_THREADS) k_msm_plan(MsmGeom g, const uint32_t* __restrict__ cnt, uint32_t* start,
                                                          uint32_t* slot_base, uint2* items, uint32_t* n_items, uint32_t* heavy,
                                                          uint32_t* n_heavy) {
    __shared__ uint32_t sh_e[PLAN_THREADS], sh_s[PLAN_THREADS];
    __shared__ uint32_t cls[MSM_CLASSES + 1];
    const uint32_t SEG = g.seg;
    // class of a slice of `len` entries: 0 for full slices, up to MSM_CLASSES - 1 for the shortest
    auto cls_of = [&](uint32_t len) { return (uint32_t)MSM_CLASSES - (uint32_t)(((uint64_t)len * MSM_CLASSES + SEG - 1) / SEG); };
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    const uint32_t* c = cnt + (size_t)b * g.n_buckets;
    uint32_t* st = start + (size_t)b * g.n_buckets;
    uint32_t* sb = slot_base + (size_t)b * (g.n_buckets + 1);
    uint2* it = items + (size_t)b * g.max_items;
    const uint32_t per = (g.n_buckets + PLAN_THREADS - 1) / PLAN_THREADS;
    const uint32_t k0 = tid * per, k1 = min(k0 + per, g.n_buckets);
    if (tid <= MSM_CLASSES) cls[tid] = 0;
    __shared__ uint32_t sh_heavy;
    if (tid == 0) sh_heavy = 0;
    uint32_t* hv = heavy + (size_t)b * g.max_heavy;
    uint32_t se = 0, ss = 0;
    for (uint32_t k = k0; k < k1; k++) {
        uint32_t v = c[k];
        se += v;
        ss += (v + SEG - 1) / SEG;
    }
    sh_e[tid] = se;
    sh_s[tid] = ss;
    __syncthreads();
    // Hillis-Steele inclusive scan over the per-thread sums
    for (int off = 1; off < PLAN_THREADS; off <<= 1) {
        uint32_t ve = 0, vs = 0;
        if (tid >= off) { ve = sh_e[tid - off]; vs = sh_s[tid - off]; }
        __syncthreads();
        sh_e[tid] += ve;
        sh_s[tid] += vs;
        __syncthreads();
    }
    uint32_t oe = sh_e[tid] - se, os = sh_s[tid] - ss;
    for (uint32_t k = k0; k < k1; k++) {
        uint32_t v = c[k];
        st[k] = oe;
        sb[k] = os;
        oe += v;
        uint32_t full = v / SEG, rem = v % SEG;
        os += full + (rem ? 1 : 0);
        if (full) atomicAdd(&cls[0], full);
        if (rem) atomicAdd(&cls[cls_of(rem)], 1u);
        if (full + (rem ? 1 : 0) >= (uint32_t)MSM_HEAVY_SEGS) hv[atomicAdd(&sh_heavy, 1u)] = k;
    }
    if (tid == PLAN_THREADS - 1) {
        sb[g.n_buckets] = sh_s[tid];
        n_items[b] = sh_s[tid];
    }
    __syncthreads();
    if (tid == 0) n_heavy[b] = sh_heavy;
    if (tid == 0) {  // exclusive scan over the length classes (class 0 = full segments first)
        uint32_t run = 0;
        for (int q = 0; q < MSM_CLASSES; q++) {
            uint32_t v = cls[q];
            cls[q] = run;
            run += v;
        }
    }
    __syncthreads();
    for (uint32_t k = k0; k < k1; k++) {
        uint32_t v = c[k];
        uint32_t full = v / SEG, rem = v % SEG;
        if (full) {
            uint32_t pos = atomicAdd(&cls[0], full);
            for (uint32_t sidx = 0; sidx < full; sidx++) it[pos + sidx] = make_uint2(k, sidx);
        }
        if (rem) {
            uint32_t pos = atomicAdd(&cls[cls_of(rem)], 1u);
            it[pos] = make_uint2(k, full);
        }
    }
}

int msm_sort(const MsmGeom& g, const uint32_t* scalars, size_t stride_words, size_t batch, const MsmSortWs& ws,
             const uint32_t* valid, cudaStream_t st) {
    if (batch == 0 || g.n_scalars == 0) return MP_OK;
    MP_CUDA_TRY(cudaMemsetAsync(ws.cnt, 0, batch * g.n_buckets * 4, st));
    MP_CUDA_TRY(cudaMemsetAsync(ws.fill, 0, batch * g.n_buckets * 4, st));
    dim3 grid(div_up(g.n_scalars, 256), (unsigned)batch);
    k_msm_hist<<<grid, 256, 0, st>>>(g, scalars, stride_words, valid, ws.cnt);
    MP_KERNEL_CHECK();
    k_msm_plan<<<(unsigned)batch, PLAN_THREADS, 0, st>>>(g, ws.cnt, ws.start, ws.slot_base, (uint2*)ws.items, ws.n_items, ws.heavy, ws.n_heavy);
    MP_KERNEL_CHECK();
    k_msm_scatter<<<grid, 256, 0, st>>>(g, scalars, stride_words, valid, ws.start, ws.fill, ws.entries);
    MP_KERNEL_CHECK();
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// bucket accumulation: one thread per work item (a bucket, or a <= g.seg slice of a large bucket)
// ---------------------------------------------------------------------------------------------------------
struct JobDev {
    MsmGeom g;
    const uint32_t *cnt, *start, *slot_base, *n_items, *entries, *heavy, *n_heavy;
    const uint2* items;
    const void* table;
    void* partial;
    void* result;
    void* scratch;
};
struct JobsArg {
    JobDev j[MSM_MAX_JOBS];
};
static JobsArg make_jobs(const MsmJob* jobs, int n) {
    JobsArg a{};
    for (int i = 0; i < n; i++) {
        a.j[i].g = jobs[i].g;
        a.j[i].cnt = jobs[i].ws.cnt;
        a.j[i].start = jobs[i].ws.start;
        a.j[i].slot_base = jobs[i].ws.slot_base;
        a.j[i].n_items = jobs[i].ws.n_items;
        a.j[i].entries = jobs[i].ws.entries;
        a.j[i].heavy = jobs[i].ws.heavy;
        a.j[i].n_heavy = jobs[i].ws.n_heavy;
        a.j[i].items = (const uint2*)jobs[i].ws.items;
        a.j[i].table = jobs[i].table;
        a.j[i].partial = jobs[i].partial;
        a.j[i].result = jobs[i].result;
        a.j[i].scratch = jobs[i].scratch;
    }
    return a;
}

template <class F, int THREADS>
__global__ void __launch_bounds__(THREADS) k_msm_accumulate(const __grid_constant__ JobsArg jobs) {
    const JobDev& J = jobs.j[blockIdx.z];
    const MsmGeom& g = J.g;
    const uint32_t b = blockIdx.y;
    const uint32_t j = blockIdx.x * THREADS + threadIdx.x;
    if (j >= J.n_items[b]) return;
    const uint2 item = J.items[(size_t)b * g.max_items + j];
    const uint32_t k = item.x, seg = item.y;
    const uint32_t total = J.cnt[(size_t)b * g.n_buckets + k];
    const uint32_t len = min(g.seg, total - seg * g.seg);
    const uint32_t* en = J.entries + (size_t)b * g.max_entries + J.start[(size_t)b * g.n_buckets + k] + seg * g.seg;
    const uint32_t slot = J.slot_base[(size_t)b * (g.n_buckets + 1) + k] + seg;
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(J.table);
    constexpr int AW = Affine<F>::WORDS;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t e = 0; e < len; e++) {
        uint32_t idx = __ldg(en + e);
        Affine<F> p = Affine<F>::load_ro(tab + (size_t)(idx & 0x7fffffffu) * AW);
        if (idx >> 31) p.y = p.y.neg();
        acc = acc.add_mixed(p);
    }
    uint32_t* out = reinterpret_cast<uint32_t*>(J.partial) + ((size_t)b * g.max_items + slot) * XYZZ<F>::WORDS;
    acc.store(out);
}

// Buckets that were split into several slices leave one partial sum per slice.  One warp folds such a bucket:
// lanes stride over its slots, then a shuffle tree adds the 32 lane sums; the total lands in the first slot and the
// others are reset to infinity, so the reduction sees a single partial per bucket.
template <class T>
MP_DEV T shfl_down_words(const T& v, int off) {
    T r;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 4); k++) dst[k] = __shfl_down_sync(0xffffffffu, src[k], off);
    return r;
}

constexpr int FOLD_WARPS_PER_BLOCK = 4, FOLD_BLOCKS = 16;
template <class F>
__global__ void __launch_bounds__(FOLD_WARPS_PER_BLOCK * 32) k_msm_fold(const __grid_constant__ JobsArg jobs) {
    const JobDev& J = jobs.j[blockIdx.z];
    const MsmGeom& g = J.g;
    const uint32_t b = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * FOLD_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const uint32_t nh = J.n_heavy[b];
    const uint32_t* hv = J.heavy + (size_t)b * g.max_heavy;
    const uint32_t* sb = J.slot_base + (size_t)b * (g.n_buckets + 1);
    XYZZ<F>* part = reinterpret_cast<XYZZ<F>*>(J.partial) + (size_t)b * g.max_items;
    for (uint32_t h = warp; h < nh; h += FOLD_BLOCKS * FOLD_WARPS_PER_BLOCK) {
        const uint32_t k = hv[h];
        const uint32_t s0 = sb[k], s1 = sb[k + 1];
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t s = s0 + lane; s < s1; s += 32) acc = acc.add(XYZZ<F>::load(part + s));
        for (int off = 16; off >= 1; off >>= 1) {
            XYZZ<F> other = shfl_down_words(acc, off);
            acc = acc.add(other);
        }
        __syncwarp();
        for (uint32_t s = s0 + lane; s < s1; s += 32) {
            if (s == s0) acc.store(part + s);          // lane 0 holds the total
            else XYZZ<F>::inf().store(part + s);
        }
    }
}

template <class F>
static int accumulate_impl(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st) {
    if (batch == 0 || n_jobs == 0) return MP_OK;
    if (n_jobs > MSM_MAX_JOBS) return MP_ERR_INVALID_ARG;
    JobsArg a = make_jobs(jobs, n_jobs);
    constexpr int THREADS = 128;
    uint32_t max_items = 0;
    for (int i = 0; i < n_jobs; i++) max_items = std::max(max_items, jobs[i].g.max_items);
    dim3 grid(div_up(max_items, THREADS), (unsigned)batch, (unsigned)n_jobs);
    k_msm_accumulate<F, THREADS><<<grid, THREADS, 0, st>>>(a);
    MP_KERNEL_CHECK();
    k_msm_fold<F><<<dim3(FOLD_BLOCKS, (unsigned)batch, (unsigned)n_jobs), FOLD_WARPS_PER_BLOCK * 32, 0, st>>>(a);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int msm_accumulate_g1(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st) { return accumulate_impl<Fq>(jobs, n_jobs, batch, st); }
int msm_accumulate_g2(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st) { return accumulate_impl<Fq2>(jobs, n_jobs, batch, st); }

// ---------------------------------------------------------------------------------------------------------
// bucket reduction  sum_k k * B_k  per window group
//   level 1: thread = chunk t of red_s1 buckets -> s_t = sum B, a_t = sum (j+1) B   (running sums; throughput work)
//   level 2: one block per (job, vector, group): with t = hi * D + lo,
//              sum_t t s_t = D * sum_hi hi * R_hi + sum_lo lo * C_lo    (R = row sums, C = column sums of s)
//            total = sum_t a_t + red_s1 * sum_t t s_t
// scratch per (vector, group): [l1pg][2] level-1 results, then [D][3] (R, C, VA), then [3] partial totals
// ---------------------------------------------------------------------------------------------------------
MP_DEV size_t red_words_per_group(const MsmGeom& g) { return max((size_t)g.l1pg * 2, (size_t)32) + (size_t)g.red_d * 3 + 3; }

template <class F>
MP_DEV XYZZ<F>* red_scratch(const JobDev& J, uint32_t b, uint32_t grp) {
    return reinterpret_cast<XYZZ<F>*>(J.scratch) + ((size_t)b * J.g.groups + grp) * red_words_per_group(J.g);
}

template <class F>
__global__ void __launch_bounds__(64) k_msm_reduce1(const __grid_constant__ JobsArg jobs) {
    const JobDev& J = jobs.j[blockIdx.z];
    const MsmGeom& g = J.g;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // chunk over all groups
    const uint32_t b = blockIdx.y;
    if (t >= g.l1pg * g.groups) return;
    const uint32_t grp = t / g.l1pg, tl = t % g.l1pg;
    const uint32_t* sb = J.slot_base + (size_t)b * (g.n_buckets + 1);
    const XYZZ<F>* part = reinterpret_cast<const XYZZ<F>*>(J.partial) + (size_t)b * g.max_items;
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    for (int j = (int)g.red_s1 - 1; j >= 0; j--) {
        uint32_t kl = tl * g.red_s1 + j;
        if (kl < g.bpg) {
            uint32_t k = grp * g.bpg + kl;
            uint32_t s0 = sb[k], s1 = sb[k + 1];
            for (uint32_t s = s0; s < s1; s++) run = run.add(XYZZ<F>::load(part + s));
        }
        acc = acc.add(run);
    }
    XYZZ<F>* sc = red_scratch<F>(J, b, grp);
    run.store(sc + (size_t)tl * 2);
    acc.store(sc + (size_t)tl * 2 + 1);
}

template <class F>
MP_DEV XYZZ<F> dbl_n(XYZZ<F> p, int n) {
    for (int i = 0; i < n; i++) p = p.dbl();
    return p;
}

// blockIdx.x = vector * groups + group, blockIdx.y = job; blockDim.x = 3 * max red_d (>= 32)
// phase 1: 3D threads: row sums R_hi, column sums C_lo of s, row sums VA_hi of a            (D serial adds)
// phase 2: weighted sums by bit decomposition: sum_q q X_q = sum_b 2^b sum_{q: bit b} X_q     (D/2 serial adds)
// phase 3: Horner over the bits (log D doublings), VA total; phase 4: combine                   (~ 2 log D + 14)
template <class F>
__global__ void __launch_bounds__(192) k_msm_reduce2(const __grid_constant__ JobsArg jobs, uint32_t batch) {
    const JobDev& J = jobs.j[blockIdx.y];
    const MsmGeom& g = J.g;
    if (blockIdx.x >= batch * g.groups) return;
    const uint32_t b = blockIdx.x / g.groups, grp = blockIdx.x % g.groups;
    const uint32_t D = g.red_d, tid = threadIdx.x;
    uint32_t LB = 0;
    while ((1u << LB) < D) LB++;
    XYZZ<F>* sc = red_scratch<F>(J, b, grp);
    XYZZ<F>* rcv = sc + max((size_t)g.l1pg * 2, (size_t)32);  // [D][3]: R, C, VA
    XYZZ<F>* tot = rcv + (size_t)D * 3;          // [3]
    // scratch for phase 2 lives behind the level-1 results that phase 1 has consumed: sc[0 .. 2*LB + 4)
    if (tid < 3 * D) {
        const uint32_t kind = tid / D, idx = tid % D;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t q = 0; q < D; q++) {
            uint32_t t = (kind == 1) ? q * D + idx : idx * D + q;   // column idx | row idx
            if (t < g.l1pg) acc = acc.add(XYZZ<F>::load(sc + (size_t)t * 2 + (kind == 2 ? 1 : 0)));
        }
        acc.store(rcv + (size_t)idx * 3 + kind);
    }
    __syncthreads();
    XYZZ<F>* bits = sc;  // [2][LB] bit sums, then [4] VA partials  (needs 2*LB + 4 <= 2 * l1pg, checked on the host)
    if (tid < 2 * LB) {
        const uint32_t w = tid / LB, bit = tid % LB;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t q = 0; q < D; q++)
            if ((q >> bit) & 1) acc = acc.add(XYZZ<F>::load(rcv + (size_t)q * 3 + w));
        acc.store(bits + tid);
    } else if (tid < 2 * LB + 4) {
        const uint32_t part = tid - 2 * LB;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t q = part; q < D; q += 4) acc = acc.add(XYZZ<F>::load(rcv + (size_t)q * 3 + 2));
        acc.store(bits + tid);
    }
    __syncthreads();
    if (tid < 2) {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (int bit = (int)LB - 1; bit >= 0; bit--) acc = acc.dbl().add(XYZZ<F>::load(bits + tid * LB + bit));
        acc.store(tot + tid);
    } else if (tid == 2) {
        XYZZ<F> acc = XYZZ<F>::load(bits + 2 * LB);
        for (int k = 1; k < 4; k++) acc = acc.add(XYZZ<F>::load(bits + 2 * LB + k));
        acc.store(tot + 2);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t ls = 0;
        while ((1u << ls) < g.red_s1) ls++;
        XYZZ<F> T = dbl_n(XYZZ<F>::load(tot), (int)LB).add(XYZZ<F>::load(tot + 1));
        T = dbl_n(T, (int)ls).add(XYZZ<F>::load(tot + 2));
        T.store(reinterpret_cast<XYZZ<F>*>(J.result) + (size_t)b * g.groups + grp);
    }
}

size_t msm_reduce_scratch_bytes(const MsmGeom& g, size_t batch, bool g2) {
    size_t words = g2 ? XYZZ<Fq2>::WORDS : XYZZ<Fq>::WORDS;
    size_t per_group = std::max<size_t>((size_t)g.l1pg * 2, 32) + (size_t)g.red_d * 3 + 3;
    return batch * g.groups * per_group * words * 4;
}

template <class F>
static int reduce_impl(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st) {
    if (batch == 0 || n_jobs == 0) return MP_OK;
    if (n_jobs > MSM_MAX_JOBS) return MP_ERR_INVALID_ARG;
    JobsArg a = make_jobs(jobs, n_jobs);
    uint32_t max_chunks = 0, max_bg = 0, max_d = 0;
    for (int i = 0; i < n_jobs; i++) {
        max_chunks = std::max(max_chunks, jobs[i].g.l1pg * (uint32_t)jobs[i].g.groups);
        max_bg = std::max(max_bg, (uint32_t)batch * (uint32_t)jobs[i].g.groups);
        max_d = std::max(max_d, jobs[i].g.red_d);
    }
    if (max_d > 64) { set_error_detail("msm reduce: level-2 grid %u exceeds 64", max_d); return MP_ERR_UNSUPPORTED; }
    k_msm_reduce1<F><<<dim3(div_up(max_chunks, 64), (unsigned)batch, (unsigned)n_jobs), 64, 0, st>>>(a);
    MP_KERNEL_CHECK();
    unsigned th = std::max(32u, 3 * max_d);
    k_msm_reduce2<F><<<dim3(max_bg, (unsigned)n_jobs), th, 0, st>>>(a, (uint32_t)batch);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int msm_reduce_g1(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st) { return reduce_impl<Fq>(jobs, n_jobs, batch, st); }
int msm_reduce_g2(const MsmJob* jobs, int n_jobs, size_t batch, cudaStream_t st) { return reduce_impl<Fq2>(jobs, n_jobs, batch, st); }

// ---------------------------------------------------------------------------------------------------------
// validity bitmaps
// ---------------------------------------------------------------------------------------------------------
template <class F>
__global__ void k_msm_validity(const uint32_t* __restrict__ bases, uint32_t n, uint32_t* bitmap, int accumulate) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool v = false;
    if (i < n) v = !Affine<F>::load(bases + (size_t)i * Affine<F>::WORDS).is_inf();
    uint32_t word = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (n + 31) / 32) {
        if (accumulate) bitmap[i >> 5] |= word;
        else bitmap[i >> 5] = word;
    }
}
int msm_validity_g1(const void* bases, uint32_t n, uint32_t* bitmap, bool accumulate, cudaStream_t st) {
    k_msm_validity<Fq><<<div_up(n, 256), 256, 0, st>>>((const uint32_t*)bases, n, bitmap, accumulate ? 1 : 0);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int msm_validity_g2(const void* bases, uint32_t n, uint32_t* bitmap, bool accumulate, cudaStream_t st) {
    k_msm_validity<Fq2><<<div_up(n, 256), 256, 0, st>>>((const uint32_t*)bases, n, bitmap, accumulate ? 1 : 0);
    MP_KERNEL_CHECK();
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// tables and Horner
// ---------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(64) k_msm_build_table(MsmGeom g, const uint32_t* __restrict__ bases, uint32_t n, uint32_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.table_stride) return;
    constexpr int AW = Affine<F>::WORDS;
    Affine<F> p = (i < n) ? Affine<F>::load(bases + (size_t)i * AW) : Affine<F>::inf();
    for (int t = 0; t < g.rows; t++) {
        p.store(out + ((size_t)t * g.table_stride + i) * AW);
        if (t + 1 < g.rows && !p.is_inf()) {
            XYZZ<F> q = XYZZ<F>::dbl_affine(p);
            q = dbl_n(q, g.c * g.groups - 1);
            p = q.to_affine();
        }
    }
}

template <class F>
__global__ void __launch_bounds__(32) k_msm_horner(MsmGeom g, const XYZZ<F>* __restrict__ res, XYZZ<F>* out, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const XYZZ<F>* r = res + (size_t)i * g.groups;
    XYZZ<F> acc = XYZZ<F>::load(r + g.groups - 1);
    for (int grp = g.groups - 2; grp >= 0; grp--) acc = dbl_n(acc, g.c).add(XYZZ<F>::load(r + grp));
    acc.store(out + i);
}

int msm_build_table_g1(const MsmGeom& g, const void* bases, uint32_t n, void* out, cudaStream_t st) {
    k_msm_build_table<Fq><<<div_up(g.table_stride, 64), 64, 0, st>>>(g, (const uint32_t*)bases, n, (uint32_t*)out);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int msm_build_table_g2(const MsmGeom& g, const void* bases, uint32_t n, void* out, cudaStream_t st) {
    k_msm_build_table<Fq2><<<div_up(g.table_stride, 64), 64, 0, st>>>(g, (const uint32_t*)bases, n, (uint32_t*)out);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int msm_horner_g1(const MsmGeom& g, const void* res, void* out, size_t count, cudaStream_t st) {
    k_msm_horner<Fq><<<div_up(count, 32), 32, 0, st>>>(g, (const XYZZ<Fq>*)res, (XYZZ<Fq>*)out, (uint32_t)count);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int msm_horner_g2(const MsmGeom& g, const void* res, void* out, size_t count, cudaStream_t st) {
    k_msm_horner<Fq2><<<div_up(count, 32), 32, 0, st>>>(g, (const XYZZ<Fq2>*)res, (XYZZ<Fq2>*)out, (uint32_t)count);
    MP_KERNEL_CHECK();
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stand-alone MSM over caller-supplied bases (no precomputed rows: one bucket set per window + Horner)
// ---------------------------------------------------------------------------------------------------------
template <class F> __global__ void k_xyzz_to_affine(const XYZZ<F>* in, Affine<F>* out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) XYZZ<F>::load(in + i).to_affine().store(out + i);
}

static int pick_window(size_t n) {
    // balance n*W mixed adds against 2*W*2^(c-1) reduction adds (each ~1.4 mixed adds)
    int best = 4;
    double best_cost = 1e300;
    for (int c = 4; c <= 16; c++) {
        int w = 255 / c + 1;
        double cost = (double)n * w + 2.8 * w * (double)(1u << (c - 1));
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

template <class F>
static int msm_standalone(int device, const uint8_t* bases, const uint64_t* scalars, size_t n, uint8_t* out_point, float* out_ms) {
    constexpr bool G2 = (FieldWords<F>::W == 24);
    const size_t pb = G2 ? MP_G2_BYTES : MP_G1_BYTES;
    if (!out_point || (n && (!bases || !scalars))) return MP_ERR_INVALID_ARG;
    if (n > (1u << 26)) { set_error_detail("msm: n = %zu exceeds 2^26", n); return MP_ERR_UNSUPPORTED; }
    MP_TRY(use_device(device));
    if (n == 0) {
        memset(out_point, 0, pb);
        out_point[pb - 1] = 0x40;
        if (out_ms) *out_ms = 0;
        return MP_OK;
    }
    MsmGeom g = msm_geom(pick_window(n), 0, (uint32_t)n, (uint32_t)n, 1);
    DevBuf d_bases, d_scalars, d_sort, d_partial, d_result, d_scratch, d_out;
    MP_TRY(d_bases.alloc(n * pb));
    MP_TRY(d_scalars.alloc(n * 32));
    MsmSortWs ws;
    MP_TRY(msm_sort_ws_alloc(ws, g, 1, d_sort));
    MP_TRY(d_partial.alloc((size_t)g.max_items * XYZZ<F>::WORDS * 4));
    MP_TRY(d_result.alloc((size_t)g.groups * XYZZ<F>::WORDS * 4));
    MP_TRY(d_scratch.alloc(msm_reduce_scratch_bytes(g, 1, G2)));
    MP_TRY(d_out.alloc(XYZZ<F>::WORDS * 4 + pb));
    MP_CUDA_TRY(cudaMemcpy(d_bases.p, bases, n * pb, cudaMemcpyHostToDevice));
    MP_CUDA_TRY(cudaMemcpy(d_scalars.p, scalars, n * 32, cudaMemcpyHostToDevice));
    if (G2) MP_TRY(points_from_ark_g2(d_bases.p, d_bases.p, n, 0));
    else MP_TRY(points_from_ark_g1(d_bases.p, d_bases.p, n, 0));
    cudaEvent_t e0, e1;
    MP_CUDA_TRY(cudaEventCreate(&e0));
    MP_CUDA_TRY(cudaEventCreate(&e1));
    MP_CUDA_TRY(cudaEventRecord(e0, 0));
    MP_TRY(msm_sort(g, d_scalars.as<uint32_t>(), n * 8, 1, ws, nullptr, 0));
    MsmJob t{};
    t.g = g;
    t.ws = ws;
    t.table = d_bases.p;
    t.partial = d_partial.p;
    t.result = d_result.p;
    t.scratch = d_scratch.p;
    if (G2) {
        MP_TRY(msm_accumulate_g2(&t, 1, 1, 0));
        MP_TRY(msm_reduce_g2(&t, 1, 1, 0));
        MP_TRY(msm_horner_g2(g, d_result.p, d_out.p, 1, 0));
    } else {
        MP_TRY(msm_accumulate_g1(&t, 1, 1, 0));
        MP_TRY(msm_reduce_g1(&t, 1, 1, 0));
        MP_TRY(msm_horner_g1(g, d_result.p, d_out.p, 1, 0));
    }
    MP_CUDA_TRY(cudaEventRecord(e1, 0));
    char* aff = d_out.as<char>() + XYZZ<F>::WORDS * 4;
    k_xyzz_to_affine<F><<<1, 1>>>((const XYZZ<F>*)d_out.p, (Affine<F>*)aff, 1);
    MP_KERNEL_CHECK();
    if (G2) MP_TRY(points_to_ark_g2(aff, aff, 1, 0));
    else MP_TRY(points_to_ark_g1(aff, aff, 1, 0));
    MP_CUDA_TRY(cudaMemcpy(out_point, aff, pb, cudaMemcpyDeviceToHost));
    float ms = 0;
    MP_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (out_ms) *out_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fixed-base helper (keygen): out[i] = k_i * G
// ---------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(64) k_fixed_base(const uint32_t* __restrict__ gen, const uint32_t* __restrict__ scalars,
                                                  uint32_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int AW = Affine<F>::WORDS;
    Affine<F> gpt = Affine<F>::load(gen);
    const uint32_t* s = scalars + i * 8;
    XYZZ<F> r = XYZZ<F>::inf();
    for (int bit = 254; bit >= 0; bit--) {
        r = r.dbl();
        if ((s[bit >> 5] >> (bit & 31)) & 1) r = r.add_mixed_cold(gpt);
    }
    r.to_affine().store(out + i * AW);
}

__global__ void k_copy_gen(uint32_t* out, int which) {
    if (which == 1) for (int c = 0; c < 2; c++) for (int k = 0; k < 12; k++) out[c * 12 + k] = G1_GEN[c][k];
    else for (int c = 0; c < 4; c++) for (int k = 0; k < 12; k++) out[c * 12 + k] = G2_GEN[c][k];
}

template <class F>
static int fixed_base_impl(int device, const uint64_t* scalars, size_t n, uint8_t* out) {
    constexpr bool G2 = (FieldWords<F>::W == 24);
    const size_t pb = G2 ? MP_G2_BYTES : MP_G1_BYTES;
    if (n && (!scalars || !out)) return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(device));
    if (n == 0) return MP_OK;
    DevBuf d_s, d_o, d_g;
    MP_TRY(d_s.alloc(n * 32));
    MP_TRY(d_o.alloc(n * pb));
    MP_TRY(d_g.alloc(pb));
    MP_CUDA_TRY(cudaMemcpy(d_s.p, scalars, n * 32, cudaMemcpyHostToDevice));
    k_copy_gen<<<1, 1>>>(d_g.as<uint32_t>(), G2 ? 2 : 1);
    MP_KERNEL_CHECK();
    k_fixed_base<F><<<div_up(n, 64), 64>>>(d_g.as<uint32_t>(), d_s.as<uint32_t>(), d_o.as<uint32_t>(), n);
    MP_KERNEL_CHECK();
    if (G2) MP_TRY(points_to_ark_g2(d_o.p, d_o.p, n, 0));
    else MP_TRY(points_to_ark_g1(d_o.p, d_o.p, n, 0));
    MP_CUDA_TRY(cudaMemcpy(out, d_o.p, n * pb, cudaMemcpyDeviceToHost));
    return MP_OK;
}

}  // namespace mp

using namespace mp;

extern "C" {

int mp_msm_g1(int device, const uint8_t* bases, const uint64_t* scalars, size_t n, uint8_t out_point[MP_G1_BYTES], float* out_ms) {
    return msm_standalone<Fq>(device, bases, scalars, n, out_point, out_ms);
}
int mp_msm_g2(int device, const uint8_t* bases, const uint64_t* scalars, size_t n, uint8_t out_point[MP_G2_BYTES], float* out_ms) {
    return msm_standalone<Fq2>(device, bases, scalars, n, out_point, out_ms);
}
int mp_fixed_base_g1(int device, const uint64_t* scalars, size_t n, uint8_t* out) { return fixed_base_impl<Fq>(device, scalars, n, out); }
int mp_fixed_base_g2(int device, const uint64_t* scalars, size_t n, uint8_t* out) { return fixed_base_impl<Fq2>(device, scalars, n, out); }

}  // extern "C"
