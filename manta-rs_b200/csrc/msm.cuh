// Internal interface of the bucket-method MSM (msm.cu) used by the stand-alone entry points and by the prover.
//
// Replaces ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul`.  Pipeline per scalar list:
//   digits+histogram -> scan/plan -> scatter (counting sort by bucket) -> bucket accumulate (XYZZ mixed adds)
//   -> 3-level running-sum bucket reduction -> (optional) Horner over window groups.
// Signed c-bit digits: 2^(c-1) buckets per window group.  A "table" holds `rows` precomputed multiples
// 2^(c*groups*t) * P_i (t < rows) so that window w = t*groups + g lands in bucket set g with point row t;
// rows == windows (groups == 1) removes the Horner step entirely (used for the circuit keys).
#pragma once
#include "common.cuh"

namespace mp {

constexpr int MSM_SEG = 64;            // max entries one thread accumulates (large buckets are split)
constexpr int MSM_RED_S1 = 32;         // buckets per thread in reduction level 1
constexpr int MSM_RED_S2 = 32;         // level-1 chunks per thread in level 2

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
    uint32_t max_items;      // n_buckets + max_entries / MSM_SEG + 1
    uint32_t l1_chunks;      // n_buckets / MSM_RED_S1 (bpg is padded to a multiple)
    uint32_t l2_chunks;      // ceil(l1_chunks_per_group / MSM_RED_S2) * groups
};
MsmGeom msm_geom(int c, int groups, uint32_t n_scalars, uint32_t table_stride);

// Per-list sort workspace for a batch of `batch` independent scalar vectors.
struct MsmSortWs {
    uint32_t* cnt = nullptr;        // [batch][n_buckets]   entries per bucket
    uint32_t* start = nullptr;      // [batch][n_buckets]   first entry of bucket
    uint32_t* slot_base = nullptr;  // [batch][n_buckets+1] first partial-sum slot of bucket
    uint32_t* fill = nullptr;       // [batch][n_buckets]   scatter cursors
    uint32_t* items = nullptr;      // [batch][max_items]   work items (bucket << 10 | segment), longest first
    uint32_t* n_items = nullptr;    // [batch]
    uint32_t* entries = nullptr;    // [batch][max_entries] sorted (sign << 31 | table index)
    size_t bytes(const MsmGeom& g, size_t batch) const;
};
int msm_sort_ws_alloc(MsmSortWs& ws, const MsmGeom& g, size_t batch, DevBuf& backing);

// scalars: [batch][scalar_stride] canonical Fr (8 x u32 each); only the first n_scalars of each row are used.
int msm_sort(const MsmGeom& g, const uint32_t* scalars, size_t scalar_stride_words, size_t batch, const MsmSortWs& ws,
             cudaStream_t st);

// One accumulate launch handles `n_msm` MSMs that share a sorted list (blockIdx.z selects the table).
struct MsmTables {
    const void* table[4];   // device tables (Affine<F>[rows][table_stride])
    void* partial[4];       // [batch][max_items] XYZZ<F> partial sums, indexed by slot
    void* result[4];        // [batch][groups] XYZZ<F> group results after reduction
    int n_msm;
};
int msm_accumulate_g1(const MsmGeom& g, const MsmSortWs& ws, const MsmTables& t, size_t batch, cudaStream_t st);
int msm_accumulate_g2(const MsmGeom& g, const MsmSortWs& ws, const MsmTables& t, size_t batch, cudaStream_t st);
// scratch: [n_msm][batch][l1_chunks + l2_chunks][3] XYZZ<F>
size_t msm_reduce_scratch_bytes(const MsmGeom& g, size_t batch, int n_msm, bool g2);
int msm_reduce_g1(const MsmGeom& g, const MsmSortWs& ws, const MsmTables& t, size_t batch, void* scratch, cudaStream_t st);
int msm_reduce_g2(const MsmGeom& g, const MsmSortWs& ws, const MsmTables& t, size_t batch, void* scratch, cudaStream_t st);

// Table construction: rows multiples of every base, 2^(c*groups*t) * P_i, affine, Montgomery.
// bases: Affine<F>[n] on the device (already Montgomery); out: Affine<F>[rows][stride]; entries >= n are infinity.
int msm_build_table_g1(const MsmGeom& g, const void* bases, uint32_t n, void* out, cudaStream_t st);
int msm_build_table_g2(const MsmGeom& g, const void* bases, uint32_t n, void* out, cudaStream_t st);

// Horner over window groups: out = sum_g 2^(c*g) * result[g]  (one thread per MSM; only when groups > 1)
int msm_horner_g1(const MsmGeom& g, const void* group_results, void* out_xyzz, size_t count, cudaStream_t st);
int msm_horner_g2(const MsmGeom& g, const void* group_results, void* out_xyzz, size_t count, cudaStream_t st);

}  // namespace mp
