/*
 * spconv_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle + timed CPU baseline).
 *
 * A plain-C restatement of the reference's *CPU* rulebook (indice-pair) algorithm.
 * Nothing in the product path (spconv_b200/) may import, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed host baseline.
 *
 * Parity status: PINNED TO THE REFERENCE ITSELF.  The reference's package cannot be built in
 * this image (pccm / cumm / ccimport absent, no network), but its CPU rulebook is plain C++ held
 * in Python f-strings; oracle/make_ref.py extracts that text from /root/reference, compiles it
 * with g++ into oracle/_ref/libspconv_ref.so, and tests/test_oracle_ref.py checks this file
 * against it BIT FOR BIT (pair order, first-touch output order, counts, error cases) on config 1,
 * on every rulebook geometry of the GPU parity tests (1-D .. 4-D, strided, dilated, transposed),
 * on the reference's LiDAR fixture coordinates and on KITTI-shaped synthetic clouds.  In
 * addition tests/test_oracle.py runs the reference's own dense-convolution equivalence
 * criterion (test/test_conv.py:247-357) and the fixture facts of BASELINE.md section 2
 * (P = 788 888, M = 136 998 for test/data/test_spconv.pkl).
 *
 * Reference locations restated here (all relative to /root/reference):
 *   spconv/csrc/sparse/indices.py:105-136   linear key = row-major over [batch, dims...];
 *                                            kernel offset <-> (r0..) row-major, last fastest
 *   spconv/csrc/sparse/indices.py:141-203   query_npq        (regular conv in -> out)
 *   spconv/csrc/sparse/indices.py:205-219   query_npq_no_stride (SubM)
 *   spconv/csrc/sparse/indices.py:253-269   query_nhw_out    (transposed conv)
 *   spconv/csrc/sparse/indices.py:1640-1708 SparseConvIndicesCPU::generate_subm_conv_inds
 *   spconv/csrc/sparse/indices.py:1711-1778 SparseConvIndicesCPU::generate_conv_inds
 *   spconv/csrc/sparse/all.py:1491-1530     get_conv_output_size / get_deconv_output_size
 *   spconv/csrc/sparse/all.py:2064-2127     buffer shapes, -1 / 0 pre-fill (done by caller)
 *
 * std::unordered_map<int, int> in the reference is used only through insert() (first
 * insertion of a key wins) and find(); the open-addressing map below has exactly those
 * two operations, so results are identical for any key set.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_NDIM 4

/* ---------------- first-insert-wins int64 -> int32 map ---------------- */
typedef struct {
    int64_t *keys;
    int32_t *vals;
    uint64_t cap_mask;
} orc_map;

static uint64_t orc_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33; return x;
}

static int orc_map_init(orc_map *m, size_t n_expected) {
    size_t cap = 16;
    while (cap < n_expected * 2 + 2) cap <<= 1;
    m->keys = (int64_t *)malloc(cap * sizeof(int64_t));
    m->vals = (int32_t *)malloc(cap * sizeof(int32_t));
    if (!m->keys || !m->vals) return -1;
    for (size_t i = 0; i < cap; ++i) m->keys[i] = INT64_MIN;
    m->cap_mask = cap - 1;
    return 0;
}

static void orc_map_free(orc_map *m) { free(m->keys); free(m->vals); }

/* returns pointer to the stored value, or NULL if absent */
static int32_t *orc_map_find(orc_map *m, int64_t key) {
    uint64_t h = orc_mix((uint64_t)key) & m->cap_mask;
    while (m->keys[h] != INT64_MIN) {
        if (m->keys[h] == key) return &m->vals[h];
        h = (h + 1) & m->cap_mask;
    }
    return NULL;
}

/* insert if absent (std::unordered_map::insert semantics: an existing key keeps its value) */
static void orc_map_insert(orc_map *m, int64_t key, int32_t val) {
    uint64_t h = orc_mix((uint64_t)key) & m->cap_mask;
    while (m->keys[h] != INT64_MIN) {
        if (m->keys[h] == key) return;
        h = (h + 1) & m->cap_mask;
    }
    m->keys[h] = key;
    m->vals[h] = val;
}

/* row-major linear key over [batch, d0, d1, ...]   (indices.py:105-110 layout_npq) */
static int64_t orc_linear(const int32_t *c, const int *dims, int ndim) {
    int64_t k = c[0];
    for (int a = 0; a < ndim; ++a) k = k * (int64_t)dims[a] + c[a + 1];
    return k;
}

/* kernel offset -> per-axis tap, row-major, last axis fastest (indices.py:117-136) */
static void orc_offset_to_taps(int k, const int *ksize, int ndim, int *r) {
    for (int a = ndim - 1; a >= 0; --a) { r[a] = k % ksize[a]; k /= ksize[a]; }
}

int orc_conv_output_size(int ndim, const int *in_dims, const int *ksize, const int *stride,
                         const int *padding, const int *dilation, int *out_dims) {
    /* all.py:1491-1509 */
    for (int i = 0; i < ndim; ++i) {
        if (ksize[i] == -1) out_dims[i] = 1;
        else out_dims[i] = (in_dims[i] + 2 * padding[i] - dilation[i] * (ksize[i] - 1) - 1) / stride[i] + 1;
    }
    return 0;
}

int orc_deconv_output_size(int ndim, const int *in_dims, const int *ksize, const int *stride,
                           const int *padding, const int *dilation, const int *out_padding,
                           int *out_dims) {
    /* all.py:1511-1530 (dilation is unused there too) */
    (void)dilation;
    for (int i = 0; i < ndim; ++i) {
        if (ksize[i] == -1) return -1;
        out_dims[i] = (in_dims[i] - 1) * stride[i] - 2 * padding[i] + ksize[i] + out_padding[i];
    }
    return 0;
}

/*
 * SubM rulebook, CPU order (indices.py:1640-1708).
 *   indices [N, ndim+1] int32 (b, d0, d1, ...)
 *   pairs   [2, kv, N]  int32, caller pre-fills with -1   (all.py:2071-2072)
 *   num     [kv]        int32, caller pre-fills with 0    (all.py:2074-2075)
 * returns N (== number of active outputs), or <0 on error (-2: even ksize).
 */
int orc_subm_rulebook(const int32_t *indices, int N, int ndim, int batch_size, const int *dims,
                      const int *ksize, const int *dilation, int32_t *pairs, int32_t *num) {
    if (ndim < 1 || ndim > ORC_MAX_NDIM) return -1;
    int kv = 1, pad[ORC_MAX_NDIM];
    for (int a = 0; a < ndim; ++a) {
        if (ksize[a] % 2 != 1) return -2;              /* "subm only support odd ksize" */
        pad[a] = (ksize[a] / 2) * dilation[a];         /* stride = 1 */
        kv *= ksize[a];
    }
    (void)batch_size;
    orc_map map;
    if (orc_map_init(&map, (size_t)N)) return -3;
    const int stride1 = ndim + 1;
    for (int i = 0; i < N; ++i)
        orc_map_insert(&map, orc_linear(indices + (size_t)i * stride1, dims, ndim), i);

    int32_t *pin = pairs, *pout = pairs + (size_t)kv * N;
    for (int k = 0; k < kv / 2 + 1; ++k) {
        size_t off = (size_t)k * N, off1 = (size_t)(kv - 1 - k) * N;
        if (k == kv / 2) {
            for (int i = 0; i < N; ++i) { pin[off + i] = i; pout[off + i] = i; }
            continue;
        }
        int r[ORC_MAX_NDIM];
        orc_offset_to_taps(k, ksize, ndim, r);
        for (int i = 0; i < N; ++i) {
            const int32_t *c = indices + (size_t)i * stride1;
            int32_t o[ORC_MAX_NDIM + 1];
            int valid = (c[0] < batch_size) && (c[0] >= 0);
            o[0] = c[0];
            for (int a = 0; a < ndim; ++a) {          /* query_npq_no_stride */
                o[a + 1] = c[a + 1] + pad[a] - r[a] * dilation[a];
                valid = valid && o[a + 1] >= 0 && o[a + 1] < dims[a];
            }
            if (!valid) continue;
            int32_t *hit = orc_map_find(&map, orc_linear(o, dims, ndim));
            if (!hit) continue;
            int32_t j = num[k]++;
            pin[off + j] = i;     pout[off + j] = *hit;
            pin[off1 + j] = *hit; pout[off1 + j] = i;
        }
    }
    orc_map_free(&map);
    return N;
}

/*
 * Regular / transposed conv rulebook, CPU order (indices.py:1711-1778).
 *   out_inds [kv*N, ndim+1] int32 scratch; first num_act rows are valid on return
 *   pairs    [2, kv, N] int32 pre-filled -1; num [kv] pre-filled 0
 * returns num_act (first-touch order, offset-major traversal), or <0 on error.
 */
int orc_conv_rulebook(const int32_t *indices, int N, int ndim, int batch_size, const int *out_dims,
                      const int *in_dims, const int *ksize, const int *stride, const int *padding,
                      const int *dilation, int transposed, int32_t *pairs, int32_t *out_inds,
                      int32_t *num) {
    if (ndim < 1 || ndim > ORC_MAX_NDIM) return -1;
    (void)in_dims;
    int kv = 1;
    for (int a = 0; a < ndim; ++a) kv *= ksize[a];
    orc_map map;
    if (orc_map_init(&map, (size_t)N * (size_t)(kv < 8 ? kv : 8))) return -3;
    /* the map may need to grow: outputs <= kv*N.  Rebuild-on-load keeps first-touch values. */
    size_t map_count = 0;
    const int stride1 = ndim + 1;
    int32_t *pin = pairs, *pout = pairs + (size_t)kv * N;
    int32_t num_act = 0;
    for (int k = 0; k < kv; ++k) {
        int r[ORC_MAX_NDIM];
        orc_offset_to_taps(k, ksize, ndim, r);
        size_t off = (size_t)k * N;
        for (int i = 0; i < N; ++i) {
            const int32_t *c = indices + (size_t)i * stride1;
            int32_t o[ORC_MAX_NDIM + 1];
            int valid = (c[0] < batch_size) && (c[0] >= 0);
            o[0] = c[0];
            if (transposed) {                              /* query_nhw_out */
                for (int a = 0; a < ndim; ++a) {
                    o[a + 1] = c[a + 1] * stride[a] - padding[a] + r[a] * dilation[a];
                    valid = valid && o[a + 1] >= 0 && o[a + 1] < out_dims[a];
                }
            } else {                                       /* query_npq */
                for (int a = 0; a < ndim; ++a) {
                    int h = c[a + 1] + padding[a] - r[a] * dilation[a];
                    o[a + 1] = h / stride[a];              /* C division, as the reference */
                    valid = valid && o[a + 1] >= 0 && o[a + 1] < out_dims[a] && !(h % stride[a]);
                }
            }
            if (!valid) continue;
            int64_t key = orc_linear(o, out_dims, ndim);
            int32_t *hit = orc_map_find(&map, key);
            int32_t hv;
            if (!hit) {
                hv = num_act++;
                if ((map_count + 1) * 2 > map.cap_mask + 1) {     /* grow x4, re-insert */
                    orc_map big;
                    if (orc_map_init(&big, (map.cap_mask + 1) * 2)) { orc_map_free(&map); return -3; }
                    for (uint64_t s = 0; s <= map.cap_mask; ++s)
                        if (map.keys[s] != INT64_MIN) orc_map_insert(&big, map.keys[s], map.vals[s]);
                    orc_map_free(&map);
                    map = big;
                }
                orc_map_insert(&map, key, hv);
                ++map_count;
                memcpy(out_inds + (size_t)hv * stride1, o, sizeof(int32_t) * stride1);
            } else {
                hv = *hit;
            }
            int32_t j = num[k]++;
            pin[off + j] = i;
            pout[off + j] = hv;
        }
    }
    orc_map_free(&map);
    return num_act;
}

/*
 * Row gather / scatter-add used by the reference CPU conv loop
 * (spconv/csrc/sparse/gather.py:30-86 GatherCPU::gather / scatter_add), fp32.
 */
void orc_gather_f32(float *buf, const float *src, const int32_t *inds, int n, int channels) {
    for (int i = 0; i < n; ++i)
        memcpy(buf + (size_t)i * channels, src + (size_t)inds[i] * channels, sizeof(float) * channels);
}

void orc_scatter_add_f32(float *dst, const float *buf, const int32_t *inds, int n, int channels) {
    for (int i = 0; i < n; ++i) {
        float *d = dst + (size_t)inds[i] * channels;
        const float *s = buf + (size_t)i * channels;
        for (int c = 0; c < channels; ++c) d[c] += s[c];
    }
}
