// ef_ops_predict.cu -- the step AROUND the tracker: the model prediction that produces the tracker's
// "model" inputs.  In the reference it is an OpenGL pass: IndexMap::combinedPredict draws every surfel of the map as a point
// sprite (elasticfusionpublic/Core/src/IndexMap.cpp:468-575, Shaders/splat.vert:19-88, Shaders/combo_splat.frag:19-67) into
// vertex / normal / colour / time textures under a depth test, and FillIn (Shaders/fill_vertex.frag:19-53,
// fill_normal.frag:19-55, fill_rgb.frag:19-37, ElasticFusion.cpp:756-760) patches its holes with the current frame.  As CUDA
// the predicted maps are born in device memory next to the tracker, so a closed tracking loop needs no GL context and no
// GL -> CUDA hand-over, and only the sensor frame (1.8 MB at 640x480) has to cross PCIe.
//
//   k_splat_prepare  one thread per pixel: empties the depth keys and evaluates the pixel's viewing ray once.
//   k_splat_depth    one thread per surfel: the vertex stage (cull tests, projection, sprite size from the four projected
//                    disc extremes), then every fragment of the sprite runs the ray / disc intersection of the fragment stage
//                    and competes with a 64-bit atomicMin on {24-bit depth | surfel index}: GL_LESS with submission order as
//                    the tie-break, i.e. the deterministic reading of what the rasteriser does.
//   k_splat_resolve  one thread per pixel: re-evaluates the winning surfel's fragment and writes the four targets
//                    (cleared to zero where nothing was drawn).
//   k_fill_*         the three fill-in passes, one thread per pixel.
//
// Rasterisation details OpenGL leaves to the implementation (which fragment centres a sprite of fractional size covers,
// the largest sprite, normalize / division rounding) are fixed here and stated again in the CPU restatement the tests
// compare against bit for bit; parity against a GL driver is unpinned (no GL in this image).  All arithmetic
// uses explicit round-to-nearest intrinsics (no FMA contraction, IEEE division and square root) so that the restatement
// can follow it operation by operation.
#include "ef_kernels.h"

namespace ef
{

namespace
{

constexpr float kMaxSprite = 64.f; // largest point sprite rasterised (GL: implementation limit GL_POINT_SIZE_RANGE)
constexpr unsigned long long kEmptyKey = 0xffffffffffffffffull;

struct SplatParams
{
    float t_inv[12]; // rows 0..2 of the inverse pose, row-major
    float cx, cy, fx, fy;
    int rows, cols;
    float max_depth, conf_threshold;
    int time, max_time, time_delta;
};

struct V3
{
    float x, y, z;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float dot(const V3 & a, const V3 & b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
__device__ __forceinline__ V3 normalize(const V3 & a)
{
    const float len = __fsqrt_rn(dot(a, a));
    return V3{dvd(a.x, len), dvd(a.y, len), dvd(a.z, len)};
}
__device__ __forceinline__ V3 cross(const V3 & a, const V3 & b)
{
    return V3{sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)), sub(mul(a.x, b.y), mul(a.y, b.x))};
}

// splat.vert:40-45 projectPointImage (x and y only)
__device__ __forceinline__ void project_image(const SplatParams & P, float x, float y, float z, float & u, float & v)
{
    u = add(dvd(mul(P.fx, x), z), P.cx);
    v = add(dvd(mul(P.fy, y), z), P.cy);
}

struct SurfelView
{
    V3 pos, nrm;     // in the camera frame of the prediction
    float conf, rad;
    float xw, yw, size;
};

// splat.vert:47-88; false = culled
__device__ __forceinline__ bool vertex_stage(const SplatParams & P, const float4 & pc, const float4 & ct, const float4 & nr, SurfelView & S)
{
    const float * t = P.t_inv;
    S.pos.x = add(add(add(mul(t[0], pc.x), mul(t[1], pc.y)), mul(t[2], pc.z)), t[3]);
    S.pos.y = add(add(add(mul(t[4], pc.x), mul(t[5], pc.y)), mul(t[6], pc.z)), t[7]);
    S.pos.z = add(add(add(mul(t[8], pc.x), mul(t[9], pc.y)), mul(t[10], pc.z)), t[11]);
    if(S.pos.z > P.max_depth || S.pos.z < 0 || pc.w < P.conf_threshold || sub((float)P.time, ct.w) > (float)P.time_delta || ct.w > (float)P.max_time)
        return false; // :51-55
    project_image(P, S.pos.x, S.pos.y, S.pos.z, S.xw, S.yw);
    // clip volume: the sprite is dropped when its centre leaves the viewport (:57, the GL point-clipping rule)
    if(!(S.xw >= 0.f && S.xw <= (float)P.cols && S.yw >= 0.f && S.yw <= (float)P.rows)) return false;
    S.conf = pc.w;
    S.rad = nr.w;
    const V3 rn{add(add(mul(t[0], nr.x), mul(t[1], nr.y)), mul(t[2], nr.z)), add(add(mul(t[4], nr.x), mul(t[5], nr.y)), mul(t[6], nr.z)),
                add(add(mul(t[8], nr.x), mul(t[9], nr.y)), mul(t[10], nr.z))};
    S.nrm = normalize(rn); // :62
    const V3 xd = normalize(V3{sub(S.nrm.y, S.nrm.z), -S.nrm.x, S.nrm.x}); // :64
    const float scale = mul(S.rad, 1.41421356f);
    const V3 x1{mul(xd.x, scale), mul(xd.y, scale), mul(xd.z, scale)};
    const V3 y1 = cross(S.nrm, x1); // :66
    float u[4], v[4];
    project_image(P, add(S.pos.x, x1.x), add(S.pos.y, x1.y), add(S.pos.z, x1.z), u[0], v[0]);
    project_image(P, add(S.pos.x, y1.x), add(S.pos.y, y1.y), add(S.pos.z, y1.z), u[1], v[1]);
    project_image(P, sub(S.pos.x, y1.x), sub(S.pos.y, y1.y), sub(S.pos.z, y1.z), u[2], v[2]);
    project_image(P, sub(S.pos.x, x1.x), sub(S.pos.y, x1.y), sub(S.pos.z, x1.z), u[3], v[3]);
    const float xmin = fminf(u[0], fminf(u[1], fminf(u[2], u[3]))), xmax = fmaxf(u[0], fmaxf(u[1], fmaxf(u[2], u[3])));
    const float ymin = fminf(v[0], fminf(v[1], fminf(v[2], v[3]))), ymax = fmaxf(v[0], fmaxf(v[1], fmaxf(v[2], v[3])));
    const float size = fmaxf(0.f, fmaxf(fabsf(sub(xmax, xmin)), fabsf(sub(ymax, ymin)))); // :80-85
    if(!(size == size)) return false;                                                     // NaN size: nothing is rasterised
    S.size = fminf(fmaxf(size, 1.f), kMaxSprite);                                         // point sizes clamp to the supported range
    return true;
}

struct Fragment
{
    float z;        // corrected_pos.z
    unsigned depth; // 24-bit window depth
};

// combo_splat.frag:39: the viewing ray through the fragment centred on (px + 0.5, py + 0.5).  It depends on the pixel only,
// so k_splat_prepare evaluates it once per pixel and call (three IEEE divisions and a square root: ~100 instructions that
// every one of the ~20 fragments of every surfel would otherwise repeat) and the fragment stage reads it back.
__device__ __forceinline__ V3 pixel_ray(const SplatParams & P, int px, int py)
{
    const float fxc = add((float)px, 0.5f), fyc = add((float)py, 0.5f);
    return normalize(V3{dvd(sub(fxc, P.cx), P.fx), dvd(sub(fyc, P.cy), P.fy), 1.f});
}

// combo_splat.frag:37-67 for the fragment with viewing ray l; false = discarded
__device__ __forceinline__ bool fragment_stage(const SplatParams & P, const SurfelView & S, const V3 & l, Fragment & F)
{
    const float k = dvd(dot(S.pos, S.nrm), dot(l, S.nrm));
    const V3 c{mul(k, l.x), mul(k, l.y), mul(k, l.z)};                                     // :41
    const V3 d{sub(c.x, S.pos.x), sub(c.y, S.pos.y), sub(c.z, S.pos.z)};
    // :44-50.  A NaN here (degenerate surfel) is discarded; the shader's `> sqrRad` test keeps it and hands the depth test a NaN
    // gl_FragDepth, whose outcome GL leaves to the implementation (it fails GL_LESS on the drivers we know of): same picture
    if(!(dot(d, d) <= mul(S.rad, S.rad))) return false;
    F.z = c.z;
    float depth = add(dvd(c.z, mul(2.f, P.max_depth)), 0.5f);                              // :66
    if(!(depth >= 0.f && depth <= 1.f)) return false;                                      // outside the depth range
    F.depth = (unsigned)__float2uint_rn(mul(depth, 16777215.f));                           // 24-bit depth buffer
    return true;
}

__device__ __forceinline__ void sprite_box(const SurfelView & S, int rows, int cols, int & x0, int & x1, int & y0, int & y1)
{
    // fragment centres (i + 0.5) inside [w - size/2, w + size/2)
    const float h = mul(S.size, 0.5f);
    x0 = max(0, (int)ceilf(sub(sub(S.xw, h), 0.5f)));
    x1 = min(cols, (int)ceilf(sub(add(S.xw, h), 0.5f)));
    y0 = max(0, (int)ceilf(sub(sub(S.yw, h), 0.5f)));
    y1 = min(rows, (int)ceilf(sub(add(S.yw, h), 0.5f)));
}

__device__ __forceinline__ void load_surfel(const float * __restrict__ surfels, size_t stride_floats, int i, float4 & pc, float4 & ct, float4 & nr)
{
    const float4 * s = reinterpret_cast<const float4 *>(surfels + (size_t)i * stride_floats);
    pc = __ldg(s);     // position, confidence
    ct = __ldg(s + 1); // colour, instance colour, init time, time
    nr = __ldg(s + 2); // normal, radius
}

__global__ void __launch_bounds__(256) k_splat_prepare(const SplatParams P, unsigned long long * __restrict__ keys, float4 * __restrict__ rays)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.rows * P.cols) return;
    keys[i] = kEmptyKey;
    const int y = i / P.cols, x = i - y * P.cols;
    const V3 l = pixel_ray(P, x, y);
    rays[i] = make_float4(l.x, l.y, l.z, 0.f);
}

__global__ void __launch_bounds__(256) k_splat_depth(const float * __restrict__ surfels, size_t stride_floats, int count, const SplatParams P,
                                                     unsigned long long * __restrict__ keys, const float4 * __restrict__ rays)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= count) return;
    float4 pc, ct, nr;
    load_surfel(surfels, stride_floats, i, pc, ct, nr);
    SurfelView S;
    if(!vertex_stage(P, pc, ct, nr, S)) return;
    int x0, x1, y0, y1;
    sprite_box(S, P.rows, P.cols, x0, x1, y0, y1);
    for(int y = y0; y < y1; y++)
        for(int x = x0; x < x1; x++)
        {
            Fragment F;
            const float4 r = __ldg(rays + (size_t)y * P.cols + x);
            if(fragment_stage(P, S, V3{r.x, r.y, r.z}, F)) atomicMin(keys + (size_t)y * P.cols + x, ((unsigned long long)F.depth << 32) | (unsigned)i);
        }
}

__global__ void __launch_bounds__(256) k_splat_resolve(const float * __restrict__ surfels, size_t stride_floats, const SplatParams P,
                                                       const unsigned long long * __restrict__ keys, const float4 * __restrict__ rays,
                                                       uchar4 * __restrict__ image,
                                                       float4 * __restrict__ vertex, float4 * __restrict__ normal, uint16_t * __restrict__ time,
                                                       uchar4 * __restrict__ inst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.rows * P.cols) return;
    const unsigned long long key = keys[i];
    uchar4 img = make_uchar4(0, 0, 0, 0), ins = img;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;
    uint16_t tm = 0;
    if(key != kEmptyKey)
    {
        const int y = i / P.cols, x = i - y * P.cols;
        float4 pc, ct, nr;
        load_surfel(surfels, stride_floats, (int)(unsigned)key, pc, ct, nr);
        SurfelView S;
        Fragment F;
        const float4 r = __ldg(rays + i);
        if(vertex_stage(P, pc, ct, nr, S) && fragment_stage(P, S, V3{r.x, r.y, r.z}, F))
        {
            const float fxc = add((float)x, 0.5f), fyc = add((float)y, 0.5f);
            const int rgb = (int)ct.x; // color.glsl decodeColor; the RGBA8 target stores the bytes back
            img = make_uchar4((rgb >> 16) & 0xff, (rgb >> 8) & 0xff, rgb & 0xff, 255);
            const int irgb = (int)ct.y; // InstanceFusion's fifth render target: inst = decodeColor(colTime.y), combo_splat.frag:29, :54
            ins = make_uchar4((irgb >> 16) & 0xff, (irgb >> 8) & 0xff, irgb & 0xff, 255);
            v = make_float4(mul(mul(sub(fxc, P.cx), F.z), dvd(1.f, P.fx)), mul(mul(sub(fyc, P.cy), F.z), dvd(1.f, P.fy)), F.z, S.conf); // :61
            n = make_float4(S.nrm.x, S.nrm.y, S.nrm.z, S.rad);
            tm = (uint16_t)(unsigned)ct.z; // :65
        }
    }
    if(image) image[i] = img;
    vertex[i] = v;
    normal[i] = n;
    if(time) time[i] = tm;
    if(inst) inst[i] = ins;
}

// ---- FillIn ----
struct FillParams
{
    float cx, cy, inv_fx, inv_fy; // FillIn.cpp passes cam = (cx, cy, 1/fx, 1/fy)
    int rows, cols, passthrough;
};

__device__ __forceinline__ V3 raw_vertex(const FillParams & P, const uint16_t * __restrict__ depth, int x, int y)
{
    // texture fetches clamp to the edge
    const int xs = min(max(x, 0), P.cols - 1), ys = min(max(y, 0), P.rows - 1);
    const float z = dvd((float)__ldg(depth + (size_t)ys * P.cols + xs), 1000.0f);
    return V3{mul(mul(sub((float)x, P.cx), z), P.inv_fx), mul(mul(sub((float)y, P.cy), z), P.inv_fy), z};
}

__global__ void __launch_bounds__(256) k_fill_vertex(const float4 * __restrict__ predicted, const uint16_t * __restrict__ depth, const FillParams P,
                                                     float4 * __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.rows * P.cols) return;
    const float4 s = __ldg(predicted + i);
    if(s.z == 0 || P.passthrough == 1) // fill_vertex.frag:44-48
    {
        const int y = i / P.cols, x = i - y * P.cols;
        const V3 v = raw_vertex(P, depth, x, y);
        out[i] = make_float4(v.x, v.y, v.z, 1.f);
    }
    else
        out[i] = s;
}

__global__ void __launch_bounds__(256) k_fill_normal(const float4 * __restrict__ predicted, const uint16_t * __restrict__ depth, const FillParams P,
                                                     float4 * __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= P.rows * P.cols) return;
    const float4 s = __ldg(predicted + i);
    if(s.z == 0 || P.passthrough == 1) // fill_normal.frag:46-51, geometry.glsl:49-59 (forward differences on the raw depth)
    {
        const int y = i / P.cols, x = i - y * P.cols;
        const V3 v = raw_vertex(P, depth, x, y), vx = raw_vertex(P, depth, x + 1, y), vy = raw_vertex(P, depth, x, y + 1);
        const V3 n = normalize(cross(V3{sub(vx.x, v.x), sub(vx.y, v.y), sub(vx.z, v.z)}, V3{sub(vy.x, v.x), sub(vy.y, v.y), sub(vy.z, v.z)}));
        out[i] = make_float4(n.x, n.y, n.z, 1.f);
    }
    else
        out[i] = s;
}

__global__ void __launch_bounds__(256) k_fill_rgb(const uchar4 * __restrict__ predicted, const uchar4 * __restrict__ raw, int n, int passthrough,
                                                  uchar4 * __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uchar4 s = __ldg(predicted + i);
    out[i] = ((int)s.x + (int)s.y + (int)s.z == 0 || passthrough == 1) ? __ldg(raw + i) : s; // fill_rgb.frag:31-36
}

} // namespace

cudaError_t launch_splat_predict(const SplatArgs & a, cudaStream_t s)
{
    SplatParams P;
    for(int i = 0; i < 12; i++) P.t_inv[i] = a.t_inv[i];
    P.cx = a.cx; P.cy = a.cy; P.fx = a.fx; P.fy = a.fy;
    P.rows = a.rows; P.cols = a.cols;
    P.max_depth = a.max_depth; P.conf_threshold = a.conf_threshold;
    P.time = a.time; P.max_time = a.max_time; P.time_delta = a.time_delta;
    const int n = a.rows * a.cols;
    unsigned long long * keys = static_cast<unsigned long long *>(a.keys);
    const size_t stride = a.stride_bytes / sizeof(float);
    float4 * rays = reinterpret_cast<float4 *>(keys + n); // scratch = keys (8 B / pixel) followed by the viewing rays (16 B / pixel)
    k_splat_prepare<<<(n + 255) / 256, 256, 0, s>>>(P, keys, rays);
    if(a.count > 0) k_splat_depth<<<(a.count + 255) / 256, 256, 0, s>>>(a.surfels, stride, a.count, P, keys, rays);
    k_splat_resolve<<<(n + 255) / 256, 256, 0, s>>>(a.surfels, stride, P, keys, rays, reinterpret_cast<uchar4 *>(a.image), reinterpret_cast<float4 *>(a.vertex),
                                                    reinterpret_cast<float4 *>(a.normal), a.time_out, reinterpret_cast<uchar4 *>(a.inst));
    return cudaGetLastError();
}

cudaError_t launch_fill_vertex(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                               float * out, cudaStream_t s)
{
    const FillParams P{cx, cy, 1.f / fx, 1.f / fy, rows, cols, passthrough};
    k_fill_vertex<<<(rows * cols + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4 *>(predicted), depth, P, reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

cudaError_t launch_fill_normal(const float * predicted, const uint16_t * depth, int rows, int cols, float cx, float cy, float fx, float fy, int passthrough,
                               float * out, cudaStream_t s)
{
    const FillParams P{cx, cy, 1.f / fx, 1.f / fy, rows, cols, passthrough};
    k_fill_normal<<<(rows * cols + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4 *>(predicted), depth, P, reinterpret_cast<float4 *>(out));
    return cudaGetLastError();
}

cudaError_t launch_fill_rgb(const uint8_t * predicted, const uint8_t * raw, int rows, int cols, int passthrough, uint8_t * out, cudaStream_t s)
{
    k_fill_rgb<<<(rows * cols + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uchar4 *>(predicted), reinterpret_cast<const uchar4 *>(raw), rows * cols,
                                                        passthrough, reinterpret_cast<uchar4 *>(out));
    return cudaGetLastError();
}

} // namespace ef
