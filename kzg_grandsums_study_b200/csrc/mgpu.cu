// Multi-GPU MSM inside ONE process, for hosts that cannot run one process per GPU: the N-API addon behind
// G1.multiExpAffine / Polynomial.multiExponentiation (reference src/polynomial/polynomial.js:1106-1115) lives in a single
// Node process.  (bench.py --gpus N uses the other form of the same split: one rank per GPU, torch.distributed / NCCL
// all-gather of the partials -- kzg_srs_msm_partial / kzg_srs_msm_host_partial / kzg_g1_partials_combine.)
//
// Device g owns the contiguous slice [first_g, first_g + count_g) of the SRS, resident with its window table from load
// time, and the matching slice of the scalars.  One MSM:
//   * one host thread per device enqueues that device's pipeline (scalars from host memory: the piecewise upload of
//     kzg_srs_msm_host_partial, hidden behind the pieces' MSMs) -- the enqueue itself costs ~0.3 ms of host time per
//     device, serial enqueueing would delay the last of eight devices by 2 ms of a 5 ms shard;
//   * every device leaves ONE extended-Jacobian partial point (128 bytes) in its own memory and records an event;
//   * device 0's stream waits for the events, pulls the G partials with 128-byte peer copies over NVLink and runs the
//     usual g1_finish (quad-lane sum, one inversion) -- 1 KiB in total at G = 8: there is no bandwidth to speak of, the
//     exchange is latency only, so no collective library is involved.
// NTT, scans and whole proofs do not shard at these sizes (replicas only).
#include <string.h>

#include <functional>
#include <thread>

#include "common.cuh"

struct kzg_mgpu {
    std::vector<kzg_ctx*> ctx;
    std::vector<kzg_srs*> srs;
    std::vector<uint64_t> first, count;
    std::vector<kzg_buf*> scalars;      // resident scalar shards (kzg_mgpu_scalars_upload)
    std::vector<void*> partial;         // 128 B on device g
    std::vector<cudaEvent_t> done;      // recorded on device g's stream after its partial
    void* gathered = nullptr;           // G x 128 B on device 0
    uint64_t n_points = 0;
    uint64_t n_scalars = 0;
    std::string err;
};

namespace {

using namespace kzg;

int mgpu_err(kzg_mgpu* m, int code, const std::string& msg) {
    if (m) m->err = msg;
    return code;
}

void shard_range(uint64_t n, uint32_t world, uint32_t rank, uint64_t* first, uint64_t* count) {
    // the first n % world devices take one point more (same split as sharded_msm.py::shard_range)
    const uint64_t base = n / world, extra = n % world;
    *first = rank * base + (rank < extra ? rank : extra);
    *count = base + (rank < extra ? 1 : 0);
}

void free_srs(kzg_mgpu* m) {
    for (size_t g = 0; g < m->srs.size(); g++)
        if (m->srs[g]) {
            kzg_srs_free(m->ctx[g], m->srs[g]);
            m->srs[g] = nullptr;
        }
    m->n_points = 0;
}
void free_scalars(kzg_mgpu* m) {
    for (size_t g = 0; g < m->scalars.size(); g++)
        if (m->scalars[g]) {
            kzg_buf_free(m->ctx[g], m->scalars[g]);
            m->scalars[g] = nullptr;
        }
    m->n_scalars = 0;
}

// run fn(g) for every device on its own host thread; first failure wins
template <class F>
int for_each_device(kzg_mgpu* m, F&& fn) {
    const size_t G = m->ctx.size();
    std::vector<int> rc(G, KZG_OK);
    if (G == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> workers;
        workers.reserve(G);
        for (size_t g = 0; g < G; g++) workers.emplace_back([&, g]() { rc[g] = fn((uint32_t)g); });
        for (auto& w : workers) w.join();
    }
    for (size_t g = 0; g < G; g++)
        if (rc[g] != KZG_OK) return mgpu_err(m, rc[g], "device " + std::to_string(m->ctx[g]->device) + ": " + kzg_last_error(m->ctx[g]));
    return KZG_OK;
}

// device 0 pulls the partials and finishes
int gather_and_finish(kzg_mgpu* m, uint8_t out_affine[64]) {
    const size_t G = m->ctx.size();
    kzg_ctx* c0 = m->ctx[0];
    DeviceGuard guard(c0);
    for (size_t g = 0; g < G; g++) {
        cudaError_t e = cudaStreamWaitEvent(c0->stream, m->done[g], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyPeerAsync((uint8_t*)m->gathered + 128 * g, c0->device, m->partial[g], m->ctx[g]->device, 128, c0->stream);
        if (e != cudaSuccess) return mgpu_err(m, KZG_ERR_CUDA, std::string("partial exchange: ") + cudaGetErrorString(e));
    }
    int r = kzg_g1_partials_combine(c0, m->gathered, (uint32_t)G, out_affine);
    if (r != KZG_OK) return mgpu_err(m, r, kzg_last_error(c0));
    return KZG_OK;
}

}  // namespace

extern "C" {

int kzg_mgpu_create(const int* devices, uint32_t n_devices, kzg_mgpu** out) {
    if (!out || (!devices && n_devices)) return KZG_ERR_ARG;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) return KZG_ERR_CUDA;  // no CPU fallback
    kzg_mgpu* m = new kzg_mgpu();
    std::vector<int> devs;
    if (n_devices == 0)
        for (int d = 0; d < visible; d++) devs.push_back(d);  // all visible devices
    else
        devs.assign(devices, devices + n_devices);
    int prev = 0;
    cudaGetDevice(&prev);
    int r = KZG_OK;
    for (int d : devs) {
        kzg_ctx* c = nullptr;
        r = kzg_ctx_create(d, nullptr, &c);
        if (r != KZG_OK) break;
        m->ctx.push_back(c);
        void* p = nullptr;
        cudaEvent_t ev = nullptr;
        cudaSetDevice(d);
        if (cudaMalloc(&p, 128) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            r = KZG_ERR_CUDA;
            break;
        }
        m->partial.push_back(p);
        m->done.push_back(ev);
        // peer access lets the 128-byte copies go directly over NVLink; without it the driver stages them
        for (kzg_ctx* other : m->ctx)
            if (other->device != d) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, d, other->device) == cudaSuccess && can) {
                    cudaDeviceEnablePeerAccess(other->device, 0);
                    cudaSetDevice(other->device);
                    cudaDeviceEnablePeerAccess(d, 0);
                    cudaSetDevice(d);
                }
            }
        cudaGetLastError();  // (peer access may already be enabled: not an error)
    }
    if (r == KZG_OK) {
        cudaSetDevice(m->ctx[0]->device);
        if (cudaMalloc(&m->gathered, 128 * m->ctx.size()) != cudaSuccess) r = KZG_ERR_CUDA;
    }
    cudaSetDevice(prev);
    m->srs.assign(m->ctx.size(), nullptr);
    m->scalars.assign(m->ctx.size(), nullptr);
    m->first.assign(m->ctx.size(), 0);
    m->count.assign(m->ctx.size(), 0);
    if (r != KZG_OK) {
        kzg_mgpu_destroy(m);
        return r;
    }
    *out = m;
    return KZG_OK;
}

int kzg_mgpu_destroy(kzg_mgpu* m) {
    if (!m) return KZG_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    free_scalars(m);
    free_srs(m);
    for (size_t g = 0; g < m->ctx.size(); g++) {
        cudaSetDevice(m->ctx[g]->device);
        cudaStreamSynchronize(m->ctx[g]->stream);
        if (g < m->partial.size()) cudaFree(m->partial[g]);
        if (g < m->done.size()) cudaEventDestroy(m->done[g]);
        if (g == 0) cudaFree(m->gathered);
    }
    for (kzg_ctx* c : m->ctx) kzg_ctx_destroy(c);
    cudaSetDevice(prev);
    delete m;
    return KZG_OK;
}

uint32_t kzg_mgpu_device_count(kzg_mgpu* m) { return m ? (uint32_t)m->ctx.size() : 0; }
kzg_ctx* kzg_mgpu_ctx(kzg_mgpu* m, uint32_t i) { return m && i < m->ctx.size() ? m->ctx[i] : nullptr; }
const char* kzg_mgpu_last_error(kzg_mgpu* m) { return m ? m->err.c_str() : "null handle"; }
uint64_t kzg_mgpu_srs_len(kzg_mgpu* m) { return m ? m->n_points : 0; }

int kzg_mgpu_shard(kzg_mgpu* m, uint32_t i, uint64_t* first, uint64_t* count) {
    if (!m || i >= m->ctx.size()) return KZG_ERR_ARG;
    if (first) *first = m->first[i];
    if (count) *count = m->count[i];
    return KZG_OK;
}

// ---- SRS: slice g resident on device g with its window table -------------------------------------------------------
static int install_srs(kzg_mgpu* m, uint64_t n_points, const std::function<int(uint32_t, uint64_t, uint64_t, kzg_srs**)>& make) {
    free_srs(m);
    free_scalars(m);
    const uint32_t G = (uint32_t)m->ctx.size();
    for (uint32_t g = 0; g < G; g++) shard_range(n_points, G, g, &m->first[g], &m->count[g]);
    int r = for_each_device(m, [&](uint32_t g) {
        kzg_srs* s = nullptr;
        int rc = make(g, m->first[g], m->count[g], &s);
        if (rc == KZG_OK && m->count[g]) rc = kzg_srs_precompute(m->ctx[g], s, 0);
        if (rc == KZG_OK) rc = kzg_ctx_sync(m->ctx[g]);
        if (rc != KZG_OK && s) {
            kzg_srs_free(m->ctx[g], s);
            s = nullptr;
        }
        m->srs[g] = s;
        return rc;
    });
    if (r != KZG_OK) {
        free_srs(m);
        return r;
    }
    m->n_points = n_points;
    return KZG_OK;
}

int kzg_mgpu_srs_generate(kzg_mgpu* m, const uint8_t tau_std[32], uint64_t n_points) {
    if (!m || !tau_std) return KZG_ERR_ARG;
    return install_srs(m, n_points, [&](uint32_t g, uint64_t first, uint64_t count, kzg_srs** out) {
        return kzg_srs_generate_range(m->ctx[g], tau_std, first, count, out);
    });
}

int kzg_mgpu_srs_from_host(kzg_mgpu* m, const uint8_t* affine, uint64_t n_points) {
    if (!m || (!affine && n_points)) return KZG_ERR_ARG;
    return install_srs(m, n_points, [&](uint32_t g, uint64_t first, uint64_t count, kzg_srs** out) {
        return kzg_srs_from_host(m->ctx[g], affine + 64 * first, count, out);
    });
}

// the first n_points of section 2 of a .ptau, sharded (prover.js:15-16,83-85; header checks of ptau_utils.js:3-24)
int kzg_mgpu_srs_load_ptau(kzg_mgpu* m, const char* path, uint64_t n_points) {
    if (!m || !path) return KZG_ERR_ARG;
    return install_srs(m, n_points, [&](uint32_t g, uint64_t first, uint64_t count, kzg_srs** out) {
        return kzg_srs_load_ptau_range(m->ctx[g], path, first, count, out, nullptr);
    });
}

// ---- MSM ------------------------------------------------------------------------------------------------------------
// scalars in HOST memory (n x 32 B standard form, n <= |SRS|): device g takes the part of its SRS slice
int kzg_mgpu_srs_msm_host(kzg_mgpu* m, const void* scalars_std_host, uint64_t n, uint8_t out_affine[64]) {
    if (!m || (!scalars_std_host && n) || !out_affine) return KZG_ERR_ARG;
    if (n > m->n_points) return mgpu_err(m, KZG_ERR_ARG, "msm: more scalars than SRS points");
    const uint8_t* host = (const uint8_t*)scalars_std_host;
    int r = for_each_device(m, [&](uint32_t g) {
        const uint64_t first = m->first[g];
        const uint64_t cnt = n > first ? (n - first < m->count[g] ? n - first : m->count[g]) : 0;
        int rc = kzg_srs_msm_host_partial(m->ctx[g], m->srs[g], 0, host + 32 * first, cnt, m->partial[g]);
        if (rc != KZG_OK) return rc;
        DeviceGuard guard(m->ctx[g]);
        return cudaEventRecord(m->done[g], m->ctx[g]->stream) == cudaSuccess ? (int)KZG_OK : (int)KZG_ERR_CUDA;
    });
    if (r != KZG_OK) return r;
    return gather_and_finish(m, out_affine);
}

// scalars made resident once (standard form), then any number of MSMs over them (the resident-input form of bench.py)
int kzg_mgpu_scalars_upload(kzg_mgpu* m, const void* scalars_std_host, uint64_t n) {
    if (!m || (!scalars_std_host && n)) return KZG_ERR_ARG;
    if (n > m->n_points) return mgpu_err(m, KZG_ERR_ARG, "msm: more scalars than SRS points");
    free_scalars(m);
    const uint8_t* host = (const uint8_t*)scalars_std_host;
    int r = for_each_device(m, [&](uint32_t g) {
        const uint64_t first = m->first[g];
        const uint64_t cnt = n > first ? (n - first < m->count[g] ? n - first : m->count[g]) : 0;
        int rc = kzg_buf_alloc(m->ctx[g], cnt, &m->scalars[g]);
        if (rc == KZG_OK && cnt) rc = kzg_buf_upload(m->ctx[g], m->scalars[g], 0, host + 32 * first, cnt);
        return rc;
    });
    if (r != KZG_OK) {
        free_scalars(m);
        return r;
    }
    m->n_scalars = n;
    return KZG_OK;
}

int kzg_mgpu_srs_msm(kzg_mgpu* m, uint8_t out_affine[64]) {
    if (!m || !out_affine) return KZG_ERR_ARG;
    int r = for_each_device(m, [&](uint32_t g) {
        if (!m->scalars[g]) return (int)KZG_ERR_ARG;
        int rc = kzg_srs_msm_partial(m->ctx[g], m->srs[g], 0, m->scalars[g], kzg_buf_len(m->scalars[g]), m->partial[g]);
        if (rc != KZG_OK) return rc;
        DeviceGuard guard(m->ctx[g]);
        return cudaEventRecord(m->done[g], m->ctx[g]->stream) == cudaSuccess ? (int)KZG_OK : (int)KZG_ERR_CUDA;
    });
    if (r != KZG_OK) return r;
    return gather_and_finish(m, out_affine);
}

// page-lock a caller-owned host buffer (a Node Buffer, a numpy array) for every device, so that the piecewise uploads of
// the host-scalar MSMs really overlap the compute; unregister before the buffer is freed
int kzg_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || !bytes) return KZG_ERR_ARG;
    return cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) == cudaSuccess ? (int)KZG_OK : (int)KZG_ERR_CUDA;
}
int kzg_host_unregister(void* ptr) {
    if (!ptr) return KZG_ERR_ARG;
    return cudaHostUnregister(ptr) == cudaSuccess ? (int)KZG_OK : (int)KZG_ERR_CUDA;
}

}  // extern "C"
