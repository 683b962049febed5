// optim.cu -- fused multi-tensor parameter updates for dopt.online (sgd.d, adam.d, amsgrad.d).
//
// Reference: the update rule is ordinary graph (online/source/dopt/online/sgd.d:28-94): per parameter tensor two
// degenerate sgemms for the scalar broadcasts, four pointwise launches and two D2D copies, each followed by a device sync.
// Here: ONE launch walks a table of tensors.  Arithmetic is done with explicit round-to-nearest mul/add/div/sqrt in the
// reference's own operation order -- no FMA contraction -- so the result is bit-identical to evaluating the reference
// graph op by op in fp32:
//   sgd      m' = m*mu + lr*g ;  w' = w - m'                                     sgd.d:57-64
//   nesterov m' = m*mu - lr*g ;  w' = (w + mu*m') - lr*g                         sgd.d:46-55
//   adam     nb1 = b1*beta1, nb2 = b2*beta2, eta = (alpha*sqrt(1-nb2))/(1-nb1)   adam.d:46-50
//            m' = beta1*m + (1-beta1)*g ; v' = beta2*v + ((1-beta2)*g)*g ;  w' = w - eta*(m'/(sqrt(v')+eps))   adam.d:52-66
//   amsgrad  adam + vhat' = max(vhat, v_old); vhat is tracked but not used by the update (amsgrad.d:63-70, survey F11)
// lr / mu / alpha / beta / eps / b1 / b2 are rank-0 DEVICE tensors like in the reference graphs.
// HBM-bound: sgd 5 words/param (r w,g,m; w w,m), adam 7, amsgrad 9.
#include "common.cuh"
#include <map>

namespace db {

struct ParamRow {
    float* w;
    const float* g;
    float* s0;
    float* s1;
    float* s2;
    int64_t n;
    int64_t chunk0;   // first global chunk of this tensor
};

static constexpr int kChunk = 4096;   // elements per CTA trip

__device__ __forceinline__ int find_row(const ParamRow* rows, int n_rows, int64_t chunk) {
    int lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (rows[mid].chunk0 <= chunk) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

template <int NESTEROV>
__global__ void __launch_bounds__(256) sgd_kernel(const ParamRow* __restrict__ rows, int n_rows, int64_t n_chunks,
                                                  const float* __restrict__ lr_p, const float* __restrict__ mu_p,
                                                  float gscale) {
    const float lr = lr_p[0], mu = mu_p[0];
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        int ri = find_row(rows, n_rows, ch);
        ParamRow r = rows[ri];
        int64_t base = (ch - r.chunk0) * kChunk;
        int64_t end = base + kChunk < r.n ? base + kChunk : r.n;
        bool vec = ((((uintptr_t)r.w | (uintptr_t)r.g | (uintptr_t)r.s0) & 15) == 0);
        auto upd = [&](float w, float g, float m, float& wo, float& mo) {
            if (gscale != 1.0f) g = __fmul_rn(g, gscale);
            if (NESTEROV) {
                mo = __fsub_rn(__fmul_rn(m, mu), __fmul_rn(lr, g));
                wo = __fsub_rn(__fadd_rn(w, __fmul_rn(mu, mo)), __fmul_rn(lr, g));
            } else {
                mo = __fadd_rn(__fmul_rn(m, mu), __fmul_rn(lr, g));
                wo = __fsub_rn(w, mo);
            }
        };
        if (vec) {
            int64_t i = base + (int64_t)threadIdx.x * 4;
            for (; i + 3 < end; i += 256 * 4) {
                float4 w = *(float4*)(r.w + i), m = *(float4*)(r.s0 + i);
                float4 g = r.g ? dbk::ld_stream((const float4*)(r.g + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 wo, mo;
                upd(w.x, g.x, m.x, wo.x, mo.x); upd(w.y, g.y, m.y, wo.y, mo.y);
                upd(w.z, g.z, m.z, wo.z, mo.z); upd(w.w, g.w, m.w, wo.w, mo.w);
                *(float4*)(r.w + i) = wo;
                *(float4*)(r.s0 + i) = mo;
            }
            // tail of the tensor (n % 4)
            if (end == r.n) {
                int64_t t = (r.n & ~(int64_t)3) + threadIdx.x;
                if (t >= base && t < r.n) {
                    float wo, mo;
                    upd(r.w[t], r.g ? r.g[t] : 0.f, r.s0[t], wo, mo);
                    r.w[t] = wo;
                    r.s0[t] = mo;
                }
            }
        } else {
            for (int64_t i = base + threadIdx.x; i < end; i += 256) {
                float wo, mo;
                upd(r.w[i], r.g ? r.g[i] : 0.f, r.s0[i], wo, mo);
                r.w[i] = wo;
                r.s0[i] = mo;
            }
        }
    }
}

// scalars: sc[0]=nb1 sc[1]=nb2 sc[2]=eta sc[3]=1-beta1 sc[4]=1-beta2 ; also advances b1, b2 in place
__global__ void adam_scalars_kernel(const float* alpha, const float* beta1, const float* beta2, float* b1, float* b2,
                                    float* sc) {
    float nb1 = __fmul_rn(b1[0], beta1[0]);
    float nb2 = __fmul_rn(b2[0], beta2[0]);
    float eta = __fdiv_rn(__fmul_rn(alpha[0], __fsqrt_rn(__fsub_rn(1.0f, nb2))), __fsub_rn(1.0f, nb1));
    sc[0] = nb1; sc[1] = nb2; sc[2] = eta;
    sc[3] = __fsub_rn(1.0f, beta1[0]);
    sc[4] = __fsub_rn(1.0f, beta2[0]);
    b1[0] = nb1;
    b2[0] = nb2;
}

template <int AMSGRAD>
__global__ void __launch_bounds__(256) adam_kernel(const ParamRow* __restrict__ rows, int n_rows, int64_t n_chunks,
                                                   const float* __restrict__ beta1_p, const float* __restrict__ beta2_p,
                                                   const float* __restrict__ eps_p, const float* __restrict__ sc,
                                                   float gscale) {
    const float beta1 = beta1_p[0], beta2 = beta2_p[0], eps = eps_p[0];
    const float eta = sc[2], omb1 = sc[3], omb2 = sc[4];
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        int ri = find_row(rows, n_rows, ch);
        ParamRow r = rows[ri];
        int64_t base = (ch - r.chunk0) * kChunk;
        int64_t end = base + kChunk < r.n ? base + kChunk : r.n;
        for (int64_t i = base + threadIdx.x; i < end; i += 256) {
            float g = r.g ? r.g[i] : 0.f;
            if (gscale != 1.0f) g = __fmul_rn(g, gscale);
            float m = r.s0[i], v = r.s1[i], w = r.w[i];
            float mo = __fadd_rn(__fmul_rn(beta1, m), __fmul_rn(omb1, g));
            float vo = __fadd_rn(__fmul_rn(beta2, v), __fmul_rn(__fmul_rn(omb2, g), g));
            if (AMSGRAD) r.s2[i] = fmaxf(r.s2[i], v);
            float wo = __fsub_rn(w, __fmul_rn(eta, __fdiv_rn(mo, __fadd_rn(__fsqrt_rn(vo), eps))));
            r.w[i] = wo;
            r.s0[i] = mo;
            r.s1[i] = vo;
        }
    }
}

namespace {
struct TableCache {
    std::vector<ParamRow> host;
    ParamRow* dev = nullptr;
    int64_t n_chunks = 0;
    float* scalars = nullptr;
};
static std::map<uint64_t, TableCache> g_tables;

static uint64_t hash_params(const dopt_b200_param* p, int n) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < sizeof(dopt_b200_param) * (size_t)n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h ^ (uint64_t)n;
}

// device copy of the tensor table; built once per distinct parameter list (weights keep their buffers across steps:
// the reference copies new values INTO the variables' own buffers, cuda/source/dopt/cuda/package.d:419-422)
static TableCache& table_for(const dopt_b200_param* params, int n, cudaStream_t s) {
    uint64_t key = hash_params(params, n);
    auto it = g_tables.find(key);
    if (it != g_tables.end()) return it->second;
    TableCache t;
    int64_t chunk = 0;
    for (int i = 0; i < n; ++i) {
        if (params[i].n <= 0) continue;
        ParamRow r{params[i].w, params[i].g, params[i].s0, params[i].s1, params[i].s2, params[i].n, chunk};
        chunk += ceil_div(params[i].n, kChunk);
        t.host.push_back(r);
    }
    t.n_chunks = chunk;
    if (!t.host.empty()) {
        DB_CUDA(cudaMalloc(&t.dev, t.host.size() * sizeof(ParamRow)));
        DB_CUDA(cudaMemcpyAsync(t.dev, t.host.data(), t.host.size() * sizeof(ParamRow), cudaMemcpyHostToDevice, s));
        DB_CUDA(cudaStreamSynchronize(s));
    }
    DB_CUDA(cudaMalloc(&t.scalars, 8 * sizeof(float)));
    return g_tables.emplace(key, std::move(t)).first->second;
}
}  // namespace

void sgd_update(const dopt_b200_param* params, int n, const float* lr, const float* mu, int nesterov, float gscale,
                cudaStream_t s) {
    for (int i = 0; i < n; ++i) DB_REQUIRE(params[i].n <= 0 || (params[i].w && params[i].s0), "sgd: null tensor");
    TableCache& t = table_for(params, n, s);
    if (t.n_chunks == 0) return;
    int grid = (int)std::min<int64_t>(t.n_chunks, (int64_t)sm_count() * 8);
    if (nesterov) sgd_kernel<1><<<grid, 256, 0, s>>>(t.dev, (int)t.host.size(), t.n_chunks, lr, mu, gscale);
    else sgd_kernel<0><<<grid, 256, 0, s>>>(t.dev, (int)t.host.size(), t.n_chunks, lr, mu, gscale);
    DB_LAUNCH_CHECK();
}

void adam_update(const dopt_b200_param* params, int n, const float* alpha, const float* beta1, const float* beta2,
                 const float* eps, float* b1, float* b2, int amsgrad, float gscale, cudaStream_t s) {
    for (int i = 0; i < n; ++i)
        DB_REQUIRE(params[i].n <= 0 || (params[i].w && params[i].s0 && params[i].s1 && (!amsgrad || params[i].s2)),
                   "adam: null tensor");
    TableCache& t = table_for(params, n, s);
    adam_scalars_kernel<<<1, 1, 0, s>>>(alpha, beta1, beta2, b1, b2, t.scalars);
    DB_LAUNCH_CHECK();
    if (t.n_chunks == 0) return;
    int grid = (int)std::min<int64_t>(t.n_chunks, (int64_t)sm_count() * 8);
    if (amsgrad) adam_kernel<1><<<grid, 256, 0, s>>>(t.dev, (int)t.host.size(), t.n_chunks, beta1, beta2, eps, t.scalars, gscale);
    else adam_kernel<0><<<grid, 256, 0, s>>>(t.dev, (int)t.host.size(), t.n_chunks, beta1, beta2, eps, t.scalars, gscale);
    DB_LAUNCH_CHECK();
}

}  // namespace db

extern "C" {
int dopt_b200_sgd_update(const dopt_b200_param* params, int n_params, const float* lr, const float* momentum,
                         int nesterov, float grad_scale, void* stream) {
    try {
        db::require_device();
        DB_REQUIRE(params && lr && momentum && n_params >= 0, "sgd_update: null argument");
        db::sgd_update(params, n_params, lr, momentum, nesterov, grad_scale, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
int dopt_b200_adam_update(const dopt_b200_param* params, int n_params, const float* alpha, const float* beta1,
                          const float* beta2, const float* eps, float* b1, float* b2, int amsgrad, float grad_scale,
                          void* stream) {
    try {
        db::require_device();
        DB_REQUIRE(params && alpha && beta1 && beta2 && eps && b1 && b2 && n_params >= 0, "adam_update: null argument");
        db::adam_update(params, n_params, alpha, beta1, beta2, eps, b1, b2, amsgrad, grad_scale, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
}
