// a9: tile-grid patch-coordinate extraction with exact OpenCV containment semantics and ordered compaction.
//
// Replaces the reference's Python loop (atlas_patch/services/extraction.py:83-103):
//   for contour: x0,y0,w,h = cv2.boundingRect(contour)
//     for y in range(y0, y0+h, step): for x in range(x0, x0+w, step):
//       keep iff  no hole has pointPolygonTest(hole, centre) > 0                      (extraction.py:76-81)
//            and  any of the 4 probes centre +- (P//2)//2 has pointPolygonTest(contour, probe) >= 0
//                                                                                     (utils/contours.py:22-38)
// cv2.pointPolygonTest(measureDist=False) on an int32 contour with integral query points is the integer
// crossing rule restated in SURVEY.md appendix A.1 (and in oracle/coords.py, verified against cv2):
// edges that cannot cross the ray are skipped after an on-vertex / on-horizontal-edge check; for the others
// the sign of an int64 cross product decides, and a zero cross product means "on the edge" (result 0).
// Because "on edge" only ORs and crossings only XOR, edges can be evaluated in any order.
//
// Kernel 1 (flags): one thread per candidate, CTAs never straddle contours; contour / hole vertices are tiled
//   through shared memory (every thread of the CTA walks the same edge -> smem broadcast, no bank conflicts);
//   warp ballots are written as keep bit-words plus a per-CTA count.
// Kernel 2 (emit): each CTA sums the counts of the CTAs before it (a few thousand ints), then a ballot/popc
//   prefix gives each kept candidate its row index => output order == reference order (stable compaction).
#include <vector>

#include "ap_internal.cuh"

namespace {

constexpr int CB = 256;          // candidates per CTA
constexpr int VTILE = 2048;      // vertices per smem tile

struct BlockDesc {   // one per CTA of the flag kernel
    int contour;     // contour index
    int first;       // first candidate (within the contour's grid) handled by this CTA
};

struct ContourDesc {
    int v_begin, v_end;   // vertex range in contour_xy
    int h_begin, h_end;   // hole index range
    int x0, y0;           // bounding-rect origin (min vertex)
    int nx, ny;           // grid size: ceil(w/step), ceil(h/step)
    long long cand_base;  // global index of this contour's first candidate
};

// state of one cv2.pointPolygonTest evaluation, accumulated edge by edge
struct PPT {
    int px, py;
    int counter;
    bool on_edge;
    __device__ __forceinline__ void edge(int v0x, int v0y, int vx, int vy) {
        const bool skip = (v0y <= py && vy <= py) || (v0y > py && vy > py) || (v0x < px && vx < px);
        if (skip) {
            if (py == vy && (px == vx || (py == v0y && ((v0x <= px && px <= vx) || (vx <= px && px <= v0x))))) on_edge = true;
            return;
        }
        long long dist = static_cast<long long>(py - v0y) * (vx - v0x) - static_cast<long long>(px - v0x) * (vy - v0y);
        if (dist == 0) on_edge = true;
        if (vy < v0y) dist = -dist;
        counter += dist > 0;
    }
    __device__ __forceinline__ int result() const { return on_edge ? 0 : ((counter & 1) ? 1 : -1); }
};

// Walk all edges of polygon [v_begin, v_end) for NP query points per thread, vertices staged through smem.
template <int NP>
__device__ __forceinline__ void walk_polygon(const int2* __restrict__ verts, int v_begin, int v_end, int2* s_v, PPT (&p)[NP]) {
    const int K = v_end - v_begin;
    if (K <= 0) return;
    int2 prev = verts[v_end - 1];  // closing edge: last -> first
    for (int t0 = 0; t0 < K; t0 += VTILE) {
        const int cnt = min(VTILE, K - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_v[i] = verts[v_begin + t0 + i];
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            const int2 v = s_v[i];
#pragma unroll
            for (int j = 0; j < NP; ++j) p[j].edge(prev.x, prev.y, v.x, v.y);
            prev = v;
        }
    }
}

__global__ void __launch_bounds__(CB)
coords_flags_kernel(const int2* __restrict__ cverts, const int2* __restrict__ hverts, const int* __restrict__ hole_offsets,
                    const ContourDesc* __restrict__ contours, const BlockDesc* __restrict__ blocks, int patch_src, int step,
                    unsigned* __restrict__ keep_words, int* __restrict__ block_counts) {
    __shared__ int2 s_v[VTILE];
    __shared__ int s_warp_cnt[CB / 32];
    const BlockDesc bd = blocks[blockIdx.x];
    const ContourDesc cd = contours[bd.contour];
    const int local = bd.first + threadIdx.x;
    const int ncand = cd.nx * cd.ny;
    const bool valid = local < ncand;
    const int gy = valid ? local / cd.nx : 0;
    const int gx = valid ? local - gy * cd.nx : 0;
    const int x = cd.x0 + gx * step, y = cd.y0 + gy * step;
    const int half = patch_src / 2;
    const int shift = half / 2;  // int(half * 0.5)
    const int cx = x + half, cy = y + half;

    bool in_hole = false;
    for (int h = cd.h_begin; h < cd.h_end; ++h) {
        PPT p[1];
        p[0] = PPT{cx, cy, 0, false};
        walk_polygon<1>(hverts, hole_offsets[h], hole_offsets[h + 1], s_v, p);
        in_hole |= p[0].result() > 0;
    }
    bool inside;
    if (shift > 0) {
        PPT p[4];
        p[0] = PPT{cx - shift, cy - shift, 0, false};
        p[1] = PPT{cx + shift, cy + shift, 0, false};
        p[2] = PPT{cx + shift, cy - shift, 0, false};
        p[3] = PPT{cx - shift, cy + shift, 0, false};
        walk_polygon<4>(cverts, cd.v_begin, cd.v_end, s_v, p);
        inside = p[0].result() >= 0 || p[1].result() >= 0 || p[2].result() >= 0 || p[3].result() >= 0;
    } else {
        PPT p[1];
        p[0] = PPT{cx, cy, 0, false};
        walk_polygon<1>(cverts, cd.v_begin, cd.v_end, s_v, p);
        inside = p[0].result() >= 0;
    }
    const bool keep = valid && inside && !in_hole;
    const unsigned word = __ballot_sync(0xffffffffu, keep);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        keep_words[blockIdx.x * (CB / 32) + warp] = word;
        s_warp_cnt[warp] = __popc(word);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        for (int i = 0; i < CB / 32; ++i) c += s_warp_cnt[i];
        block_counts[blockIdx.x] = c;
    }
}

__global__ void __launch_bounds__(CB)
coords_emit_kernel(const ContourDesc* __restrict__ contours, const BlockDesc* __restrict__ blocks,
                   const unsigned* __restrict__ keep_words, const int* __restrict__ block_counts, int step, int read_w,
                   int read_h, int level, int* __restrict__ out_rows, long long capacity, long long* __restrict__ out_count) {
    __shared__ long long s_red[CB / 32];
    __shared__ long long s_base;
    // exclusive prefix of the counts of all CTAs before this one
    long long part = 0;
    for (int i = threadIdx.x; i < blockIdx.x; i += CB) part += block_counts[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = 0;
        for (int i = 0; i < CB / 32; ++i) b += s_red[i];
        s_base = b;
        if (blockIdx.x == gridDim.x - 1) *out_count = b + block_counts[blockIdx.x];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long base = s_base;
    for (int w = 0; w < warp; ++w) base += __popc(keep_words[blockIdx.x * (CB / 32) + w]);
    const unsigned word = keep_words[blockIdx.x * (CB / 32) + warp];
    if (!((word >> lane) & 1u)) return;
    const long long idx = base + __popc(word & ((1u << lane) - 1u));
    if (idx >= capacity) return;
    const BlockDesc bd = blocks[blockIdx.x];
    const ContourDesc cd = contours[bd.contour];
    const int local = bd.first + threadIdx.x;
    const int gy = local / cd.nx, gx = local - gy * cd.nx;
    int* r = out_rows + idx * 5;
    r[0] = cd.x0 + gx * step;
    r[1] = cd.y0 + gy * step;
    r[2] = read_w;
    r[3] = read_h;
    r[4] = level;
}

struct HostPlan {
    std::vector<ContourDesc> contours;
    std::vector<BlockDesc> blocks;
    long long total_candidates = 0;
};

// Bounding rect and grid per contour: cv2.boundingRect of an int32 point set = (min, max-min+1)
// (extraction.py:94-97); range(x0, x0+w, step) has ceil(w/step) elements.
bool build_plan(const int32_t* cxy, const int32_t* coff, int n_contours, const int32_t* hole_first, int step, HostPlan& plan) {
    plan.contours.resize(n_contours);
    long long base = 0;
    for (int c = 0; c < n_contours; ++c) {
        ContourDesc& cd = plan.contours[c];
        cd.v_begin = coff[c];
        cd.v_end = coff[c + 1];
        cd.h_begin = hole_first ? hole_first[c] : 0;
        cd.h_end = hole_first ? hole_first[c + 1] : 0;
        cd.cand_base = base;
        if (cd.v_end <= cd.v_begin) {
            cd.x0 = cd.y0 = cd.nx = cd.ny = 0;
            continue;
        }
        int xmin = cxy[2 * cd.v_begin], xmax = xmin, ymin = cxy[2 * cd.v_begin + 1], ymax = ymin;
        for (int i = cd.v_begin + 1; i < cd.v_end; ++i) {
            const int x = cxy[2 * i], y = cxy[2 * i + 1];
            xmin = x < xmin ? x : xmin; xmax = x > xmax ? x : xmax;
            ymin = y < ymin ? y : ymin; ymax = y > ymax ? y : ymax;
        }
        cd.x0 = xmin; cd.y0 = ymin;
        const long long w = (long long)xmax - xmin + 1, h = (long long)ymax - ymin + 1;
        const long long nx = (w + step - 1) / step, ny = (h + step - 1) / step;
        if (nx * ny > 0x7fffffffLL) return false;
        cd.nx = (int)nx; cd.ny = (int)ny;
        const long long nc = nx * ny;
        for (long long f = 0; f < nc; f += CB) plan.blocks.push_back(BlockDesc{c, (int)f});
        base += nc;
    }
    plan.total_candidates = base;
    return true;
}

}  // namespace

extern "C" int64_t ap_coords_capacity(const int32_t* contour_xy, const int32_t* contour_offsets, int n_contours, int step_src) {
    if (n_contours < 0 || step_src <= 0 || (n_contours > 0 && (!contour_xy || !contour_offsets))) return AP_EINVAL;
    HostPlan plan;
    if (!build_plan(contour_xy, contour_offsets, n_contours, nullptr, step_src, plan)) return AP_EINVAL;
    return plan.total_candidates;
}

extern "C" int ap_extract_coords(ap_ctx* ctx, const int32_t* contour_xy, const int32_t* contour_offsets, int n_contours,
                                 const int32_t* hole_xy, const int32_t* hole_offsets, const int32_t* hole_first,
                                 int patch_src, int step_src, int read_w, int read_h, int level, int32_t* out_rows_dev,
                                 int32_t* out_rows_host, int64_t capacity, int64_t* out_count, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_REQUIRE(ctx, out_count != nullptr, "extract_coords: out_count is NULL");
    *out_count = 0;
    AP_REQUIRE(ctx, n_contours >= 0 && patch_src > 0 && step_src > 0, "extract_coords: bad arguments (n_contours %d patch %d step %d)",
               n_contours, patch_src, step_src);
    if (n_contours == 0) return AP_OK;
    AP_REQUIRE(ctx, contour_xy && contour_offsets && hole_first && hole_offsets, "extract_coords: NULL contour arrays");
    HostPlan plan;
    AP_REQUIRE(ctx, build_plan(contour_xy, contour_offsets, n_contours, hole_first, step_src, plan),
               "extract_coords: candidate grid of a contour exceeds 2^31");
    if (plan.blocks.empty()) return AP_OK;
    AP_REQUIRE(ctx, capacity >= 0 && (capacity == 0 || out_rows_dev || out_rows_host), "extract_coords: no output buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_holes = hole_first[n_contours];
    const int n_cv = contour_offsets[n_contours], n_hv = n_holes > 0 ? hole_offsets[n_holes] : 0;
    const int nb = (int)plan.blocks.size();

    // one scratch allocation: [cverts | hverts | hole_offsets | contours | blocks | keep_words | block_counts | count | rows?]
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t o_cv = 0, o_hv = o_cv + al((size_t)n_cv * 8), o_ho = o_hv + al((size_t)n_hv * 8 + 8),
           o_cd = o_ho + al((size_t)(n_holes + 1) * 4), o_bd = o_cd + al(plan.contours.size() * sizeof(ContourDesc)),
           o_kw = o_bd + al((size_t)nb * sizeof(BlockDesc)), o_bc = o_kw + al((size_t)nb * (CB / 32) * 4),
           o_cnt = o_bc + al((size_t)nb * 4), o_rows = o_cnt + 256;
    const bool own_rows = (out_rows_dev == nullptr);
    size_t total = o_rows + (own_rows ? al((size_t)capacity * 20) : 0);
    uint8_t* scratch = nullptr;
    AP_CHECK_CUDA(ctx, cudaMallocAsync((void**)&scratch, total, st));
    int rc = AP_OK;
    auto fail = [&](int code) { cudaFreeAsync(scratch, st); return code; };
#define AP_TRY(call)                                                                                        \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return fail(ap_set_error(ctx, AP_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)));      \
    } while (0)
    AP_TRY(cudaMemcpyAsync(scratch + o_cv, contour_xy, (size_t)n_cv * 8, cudaMemcpyHostToDevice, st));
    if (n_hv) AP_TRY(cudaMemcpyAsync(scratch + o_hv, hole_xy, (size_t)n_hv * 8, cudaMemcpyHostToDevice, st));
    AP_TRY(cudaMemcpyAsync(scratch + o_ho, hole_offsets, (size_t)(n_holes + 1) * 4, cudaMemcpyHostToDevice, st));
    AP_TRY(cudaMemcpyAsync(scratch + o_cd, plan.contours.data(), plan.contours.size() * sizeof(ContourDesc), cudaMemcpyHostToDevice, st));
    AP_TRY(cudaMemcpyAsync(scratch + o_bd, plan.blocks.data(), (size_t)nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    int32_t* rows_dev = own_rows ? reinterpret_cast<int32_t*>(scratch + o_rows) : out_rows_dev;

    ProfScope prof(ctx, st, AP_K_COORDS);
    coords_flags_kernel<<<nb, CB, 0, st>>>(reinterpret_cast<const int2*>(scratch + o_cv), reinterpret_cast<const int2*>(scratch + o_hv),
                                           reinterpret_cast<const int*>(scratch + o_ho),
                                           reinterpret_cast<const ContourDesc*>(scratch + o_cd),
                                           reinterpret_cast<const BlockDesc*>(scratch + o_bd), patch_src, step_src,
                                           reinterpret_cast<unsigned*>(scratch + o_kw), reinterpret_cast<int*>(scratch + o_bc));
    ctx->launches.fetch_add(1);
    AP_TRY(cudaGetLastError());
    coords_emit_kernel<<<nb, CB, 0, st>>>(reinterpret_cast<const ContourDesc*>(scratch + o_cd),
                                          reinterpret_cast<const BlockDesc*>(scratch + o_bd),
                                          reinterpret_cast<const unsigned*>(scratch + o_kw),
                                          reinterpret_cast<const int*>(scratch + o_bc), step_src, read_w, read_h, level, rows_dev,
                                          (long long)capacity, reinterpret_cast<long long*>(scratch + o_cnt));
    ctx->launches.fetch_add(1);
    AP_TRY(cudaGetLastError());
    long long count = 0;
    AP_TRY(cudaMemcpyAsync(&count, scratch + o_cnt, 8, cudaMemcpyDeviceToHost, st));
    AP_TRY(cudaStreamSynchronize(st));
    *out_count = count;
    if (count > capacity) rc = ap_set_error(ctx, AP_ECAPACITY, "extract_coords: %lld rows but capacity is %lld", count, (long long)capacity);
    if (rc == AP_OK && out_rows_host && count > 0) {
        AP_TRY(cudaMemcpyAsync(out_rows_host, rows_dev, (size_t)count * 20, cudaMemcpyDeviceToHost, st));
        AP_TRY(cudaStreamSynchronize(st));
    }
#undef AP_TRY
    cudaFreeAsync(scratch, st);
    return rc;
}
