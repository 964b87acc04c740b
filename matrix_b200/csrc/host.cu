// host.cu -- the host-buffer face of the C ABI (ppx_cuda_host_*): what a ppx user holding
// std::vector<ppx::MatrixS<M,N>> calls.  Inputs and outputs are HOST arrays of dense reference
// records (AoS).  The batch is cut into chunks; on each device a few streams rotate over them so that
// the H2D copy of one chunk, the kernel of the previous and the D2H copy of the one before overlap
// (PCIe is full duplex and, at ~55 GB/s against 6.5 TB/s of HBM, the only thing that matters end to end).
//
// Several devices (ppx_cuda_host_set_devices / PPX_CUDA_DEVICES): one worker thread per device, and the
// chunks are handed out from ONE shared queue, a chunk at a time, to whichever device has a free slot.
// Instances are independent, so which device computes a chunk cannot change a bit of the result
// (G-device output == 1-device output, bitwise); the queue instead of fixed ceil(n/G) blocks
// is there because the PCIe rate the devices of one host get under load is uneven (shared root ports,
// host memory on one socket): with fixed blocks the call ends when the slowest device does, with the
// queue every link stays busy until the batch is empty.
//
// Streams and device scratch are kept between calls (per device: a few non-blocking streams and a memory pool that
// retains up to PPX_HOST_POOL_MIB, default 1024, of freed scratch), so a small call costs tens of microseconds instead
// of the ~0.7 ms that creating a stream and mapping fresh device memory took; ppx_cuda_host_trim() gives it back.
// After a failed call the streams it used are destroyed, not recycled: a later call starts from clean ones.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "ppx_launch.cuh"

namespace
{

struct HostIn
{
    const void *ptr;
    size_t bytes; // per instance
};
struct HostOut
{
    void *ptr; // may be NULL (optional output)
    size_t bytes;
};

constexpr int MAX_SLOTS = 8;
constexpr int MAX_DEVICES = 64;
// defaults: 3 chunks in flight per device, 48 MiB of traffic on the widest side per chunk; PPX_HOST_SLOTS /
// PPX_HOST_CHUNK_MIB override them (tuning knobs, read once)
inline int env_slots()
{
    static const int v = [] {
        const char *e = std::getenv("PPX_HOST_SLOTS");
        const int n = e ? std::atoi(e) : 3;
        return n < 1 ? 1 : (n > MAX_SLOTS ? MAX_SLOTS : n);
    }();
    return v;
}
inline size_t env_chunk_bytes()
{
    static const size_t v = [] {
        const char *e = std::getenv("PPX_HOST_CHUNK_MIB");
        const long n = e ? std::atol(e) : 48;
        return (size_t)(n < 1 ? 1 : n) << 20;
    }();
    return v;
}

// ---- the devices the host face spreads a batch over ------------------------------------------------------------
std::mutex g_dev_mu;
std::vector<int> g_devices;  // empty: the calling thread's current device only
bool g_devices_set = false;  // false: PPX_CUDA_DEVICES has not been looked at yet

int parse_device_env(std::vector<int> &out)
{
    const char *e = std::getenv("PPX_CUDA_DEVICES");
    out.clear();
    if (!e || !*e)
        return PPX_OK;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess)
        return PPX_ERR_NO_DEVICE;
    if (std::strcmp(e, "all") == 0)
    {
        for (int d = 0; d < have && d < MAX_DEVICES; d++)
            out.push_back(d);
        return PPX_OK;
    }
    const char *p = e;
    while (*p)
    {
        char *endp = nullptr;
        const long d = std::strtol(p, &endp, 10);
        if (endp == p || d < 0 || d >= have || std::find(out.begin(), out.end(), (int)d) != out.end())
        {
            out.clear();
            return PPX_ERR_BAD_ARGUMENT;
        }
        out.push_back((int)d);
        p = (*endp == ',') ? endp + 1 : endp;
        if (*endp != ',' && *endp != 0)
        {
            out.clear();
            return PPX_ERR_BAD_ARGUMENT;
        }
    }
    return PPX_OK;
}

std::vector<int> host_devices()
{
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_devices_set)
    {
        parse_device_env(g_devices); // a malformed variable leaves the list empty (current device)
        g_devices_set = true;
    }
    return g_devices;
}

// per-device resources kept between calls
struct DevCtx
{
    std::vector<cudaStream_t> idle; // streams not in use by a call
};
std::mutex g_host_mu;
DevCtx g_host_ctx[MAX_DEVICES];

// -> `want` streams and the scratch pool of the current device
int host_acquire(int want, cudaStream_t *st, cudaMemPool_t &pool, int &device)
{
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess)
        return (int)e;
    if (device < 0 || device >= MAX_DEVICES)
        return PPX_ERR_NO_DEVICE;
    const int prc = ppx_scratch_pool(device, &pool);
    if (prc != PPX_OK)
        return prc;
    std::lock_guard<std::mutex> lock(g_host_mu);
    DevCtx &c = g_host_ctx[device];
    for (int s = 0; s < want; s++)
    {
        if (!c.idle.empty())
        {
            st[s] = c.idle.back();
            c.idle.pop_back();
        }
        else
        {
            e = cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking);
            if (e != cudaSuccess)
            {
                for (int k = 0; k < s; k++)
                    c.idle.push_back(st[k]);
                return (int)e;
            }
        }
    }
    return PPX_OK;
}
// healthy streams go back to the idle list; after a failure they are destroyed so that nothing a faulted call
// touched is handed to the next one
void host_release(int device, int n, cudaStream_t *st, bool healthy)
{
    if (!healthy)
    {
        for (int s = 0; s < n; s++)
            if (st[s])
                cudaStreamDestroy(st[s]);
        (void)cudaGetLastError();
        return;
    }
    std::lock_guard<std::mutex> lock(g_host_mu);
    for (int s = 0; s < n; s++)
        if (st[s])
            g_host_ctx[device].idle.push_back(st[s]);
}

#define PPX_TRY(expr)                 \
    do                                \
    {                                 \
        cudaError_t e_ = (expr);      \
        if (e_ != cudaSuccess)        \
        {                             \
            rc = (int)e_;             \
            goto done;                \
        }                             \
    } while (0)

// what one call shares between its device workers
struct Job
{
    const std::vector<HostIn> &ins;
    const std::vector<HostOut> &outs;
    size_t n, chunk, nchunks;
    std::atomic<size_t> next{0}; // next chunk nobody has taken yet
    std::atomic<int> rc{PPX_OK}; // first failure of any worker
    Job(const std::vector<HostIn> &i, const std::vector<HostOut> &o, size_t n_, size_t chunk_)
        : ins(i), outs(o), n(n_), chunk(chunk_), nchunks((n_ + chunk_ - 1) / chunk_)
    {
    }
    void fail(int code)
    {
        int ok = PPX_OK;
        rc.compare_exchange_strong(ok, code);
    }
};

// One device's share of a call, on the calling thread's current device: takes chunks from the job's queue until it is
// empty.  run(count, d_in[], d_out[], stream) launches the AoS kernel for one chunk.  With `shared` (several
// devices) a slot is only refilled once its previous chunk has completely finished, so that a device never holds
// more than its `nslots` chunks and the rest of the queue stays available to the others; a single device needs no
// such wait (stream order does it) and keeps the queue of the copy engines full.
template <class Run>
void device_worker(Job &job, Run &run, bool shared)
{
    const std::vector<HostIn> &ins = job.ins;
    const std::vector<HostOut> &outs = job.outs;
    int rc = PPX_OK;
    cudaStream_t st[MAX_SLOTS] = {};
    const int nslots = (int)std::min<size_t>(env_slots(), job.nchunks);
    std::vector<void *> d_in(nslots * ins.size(), nullptr), d_out(nslots * outs.size(), nullptr);
    cudaMemPool_t pool = nullptr;
    int device = 0;
    rc = host_acquire(nslots, st, pool, device);
    if (rc != PPX_OK)
    {
        job.fail(rc);
        return;
    }
    for (size_t k = 0;; k++)
    {
        const int s = (int)(k % nslots);
        if (k < (size_t)nslots)
        {
            for (size_t a = 0; a < ins.size(); a++)
                PPX_TRY(cudaMallocFromPoolAsync(&d_in[s * ins.size() + a], job.chunk * ins[a].bytes, pool, st[s]));
            for (size_t a = 0; a < outs.size(); a++)
                if (outs[a].ptr)
                    PPX_TRY(cudaMallocFromPoolAsync(&d_out[s * outs.size() + a], job.chunk * outs[a].bytes, pool, st[s]));
        }
        else if (shared)
            PPX_TRY(cudaStreamSynchronize(st[s]));
        if (job.rc.load(std::memory_order_relaxed) != PPX_OK)
            break; // another device failed: stop feeding
        const size_t c = job.next.fetch_add(1, std::memory_order_relaxed);
        if (c >= job.nchunks)
            break;
        const size_t first = c * job.chunk;
        const size_t cnt = std::min(job.chunk, job.n - first);
        for (size_t a = 0; a < ins.size(); a++)
            PPX_TRY(cudaMemcpyAsync(d_in[s * ins.size() + a], (const char *)ins[a].ptr + first * ins[a].bytes,
                                    cnt * ins[a].bytes, cudaMemcpyHostToDevice, st[s]));
        rc = run(cnt, &d_in[s * ins.size()], &d_out[s * outs.size()], (void *)st[s]);
        if (rc != PPX_OK)
            goto done;
        for (size_t a = 0; a < outs.size(); a++)
            if (outs[a].ptr)
                PPX_TRY(cudaMemcpyAsync((char *)outs[a].ptr + first * outs[a].bytes, d_out[s * outs.size() + a],
                                        cnt * outs[a].bytes, cudaMemcpyDeviceToHost, st[s]));
    }
done:
    for (int s = 0; s < nslots; s++)
    {
        for (size_t a = 0; a < ins.size(); a++)
            if (d_in[s * ins.size() + a])
            {
                const cudaError_t e = cudaFreeAsync(d_in[s * ins.size() + a], st[s]);
                if (rc == PPX_OK && e != cudaSuccess)
                    rc = (int)e;
            }
        for (size_t a = 0; a < outs.size(); a++)
            if (d_out[s * outs.size() + a])
            {
                const cudaError_t e = cudaFreeAsync(d_out[s * outs.size() + a], st[s]);
                if (rc == PPX_OK && e != cudaSuccess)
                    rc = (int)e;
            }
        const cudaError_t e = cudaStreamSynchronize(st[s]);
        if (rc == PPX_OK && e != cudaSuccess)
            rc = (int)e;
    }
    if (rc != PPX_OK)
    {
        (void)cudaGetLastError();
        job.fail(rc);
    }
    host_release(device, nslots, st, rc == PPX_OK);
}

// chunk_mult: ops whose kernels have a long per-launch tail (the persistent trust-region / Newton-loop kernels: the
// last stalled instance of a launch ends ~9 ms after the queue has drained) ask for larger chunks
template <class Run>
int pipeline_impl(const std::vector<HostIn> &ins, const std::vector<HostOut> &outs, size_t n, Run &run, size_t chunk_mult)
{
    if (n == 0)
        return PPX_OK;
    for (const HostIn &a : ins)
        if (a.ptr == nullptr)
            return PPX_ERR_BAD_ARGUMENT;
    size_t in_b = 0, out_b = 0;
    for (const HostIn &a : ins)
        in_b += a.bytes;
    for (const HostOut &a : outs)
        out_b += a.ptr ? a.bytes : 0;
    const size_t widest = std::max<size_t>(std::max(in_b, out_b), 1);
    size_t chunk = std::max<size_t>(env_chunk_bytes() * chunk_mult / widest, 1024);
    chunk = (chunk + 1023) / 1024 * 1024;
    chunk = std::min(chunk, n);

    Job job(ins, outs, n, chunk);
    const std::vector<int> devs = host_devices();
    if (devs.empty())
    {
        device_worker(job, run, false); // the calling thread's current device
        return job.rc.load();
    }
    int home = -1;
    cudaError_t e = cudaGetDevice(&home);
    if (e != cudaSuccess)
        return (int)e;
    // one worker per configured device, never more workers than chunks; the calling thread is the first of them
    const size_t workers = std::min<size_t>(devs.size(), job.nchunks);
    std::vector<std::thread> th;
    try
    {
        th.reserve(workers);
        for (size_t w = 1; w < workers; w++)
            th.emplace_back([&job, &run, &devs, w] {
                try
                {
                    const cudaError_t se = cudaSetDevice(devs[w]);
                    if (se != cudaSuccess)
                        job.fail((int)se);
                    else
                        device_worker(job, run, true);
                }
                catch (...) // nothing may leave a worker thread
                {
                    job.fail(PPX_ERR_HOST);
                }
            });
    }
    catch (...) // a worker could not be started: the call fails, the workers already running stop at their next chunk
    {
        job.fail(PPX_ERR_HOST);
    }
    e = cudaSetDevice(devs[0]);
    if (e != cudaSuccess)
        job.fail((int)e);
    else if (job.rc.load() == PPX_OK)
        device_worker(job, run, workers > 1);
    for (std::thread &t : th)
        t.join();
    cudaSetDevice(home);
    return job.rc.load();
}

// C++ exceptions (std::bad_alloc, std::system_error) stop at the C boundary
template <class Run>
int pipeline(const std::vector<HostIn> &ins, const std::vector<HostOut> &outs, size_t n, Run run, size_t chunk_mult = 1)
{
    try
    {
        return pipeline_impl(ins, outs, n, run, chunk_mult);
    }
    catch (...)
    {
        return PPX_ERR_HOST;
    }
}

// the copy loop behind ppx_cuda_host_copy_probe / ppx_cuda_host_tune_devices: all `devs` copy at once
int probe_devices(const std::vector<int> &devs, void *host, size_t bytes, int direction, double seconds, double *gbs)
{
    int home = -1;
    cudaError_t e = cudaGetDevice(&home);
    if (e != cudaSuccess)
        return (int)e;
    const size_t G = devs.size();
    const size_t slice = bytes / G / 4096 * 4096;
    const size_t piece = std::min(env_chunk_bytes(), slice / 2 / 4096 * 4096);
    if (piece == 0)
        return PPX_ERR_BAD_ARGUMENT;
    std::atomic<int> ready{0}, rc{PPX_OK};
    auto work = [&](size_t w) {
        cudaStream_t st[2] = {};
        char *d[2] = {};
        int err = PPX_OK;
        auto fail = [&](cudaError_t ce) {
            if (ce != cudaSuccess && err == PPX_OK)
                err = (int)ce;
        };
        fail(cudaSetDevice(devs[w]));
        for (int k = 0; k < 2 && err == PPX_OK; k++)
        {
            fail(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
            fail(cudaMalloc((void **)&d[k], piece));
        }
        ready.fetch_add(1);
        while (ready.load() < (int)G) // aligned start
            std::this_thread::yield();
        char *h = (char *)host + w * slice;
        const size_t half = slice / 2;
        size_t moved = 0;
        const auto t0 = std::chrono::steady_clock::now();
        double dt = 0.0;
        size_t off = 0;
        while (err == PPX_OK)
        {
            for (int k = 0; k < 2 && err == PPX_OK; k++)
            {
                // stream k walks half k of the slice; direction 2: stream 0 reads the host (H2D), stream 1 writes it
                const bool to_host = direction == 1 || (direction == 2 && k == 1);
                char *hp = h + k * half + off;
                fail(to_host ? cudaMemcpyAsync(hp, d[k], piece, cudaMemcpyDeviceToHost, st[k])
                             : cudaMemcpyAsync(d[k], hp, piece, cudaMemcpyHostToDevice, st[k]));
            }
            for (int k = 0; k < 2; k++)
                fail(cudaStreamSynchronize(st[k]));
            moved += 2 * piece;
            off += piece;
            if (off + piece > half)
                off = 0;
            dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt >= seconds)
                break;
        }
        gbs[w] = dt > 0.0 ? (double)moved / dt * 1e-9 : 0.0;
        for (int k = 0; k < 2; k++)
        {
            if (d[k])
                cudaFree(d[k]);
            if (st[k])
                cudaStreamDestroy(st[k]);
        }
        if (err != PPX_OK)
        {
            int ok = PPX_OK;
            rc.compare_exchange_strong(ok, err);
        }
    };
    std::vector<std::thread> th;
    try
    {
        th.reserve(G);
        for (size_t w = 1; w < G; w++)
            th.emplace_back(work, w);
        work(0);
    }
    catch (...) // a worker could not be started: release the ones spinning on the aligned start and report it
    {
        ready.fetch_add((int)G);
        int ok = PPX_OK;
        rc.compare_exchange_strong(ok, PPX_ERR_HOST);
    }
    for (std::thread &t : th)
        t.join();
    cudaSetDevice(home);
    return rc.load();
}

constexpr size_t D = sizeof(double);
typedef const double *cd;
typedef double *md;

} // namespace

extern "C"
{

int ppx_cuda_host_trim(void) { return ppx_scratch_trim(); }

int ppx_cuda_host_set_devices(const int *devices, int count)
try
{
    std::vector<int> v;
    if (count < 0) // all visible devices
    {
        int have = 0;
        if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1)
            return PPX_ERR_NO_DEVICE;
        for (int d = 0; d < have && d < MAX_DEVICES; d++)
            v.push_back(d);
    }
    else if (count > 0)
    {
        if (!devices || count > MAX_DEVICES)
            return PPX_ERR_BAD_ARGUMENT;
        int have = 0;
        if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1)
            return PPX_ERR_NO_DEVICE;
        for (int i = 0; i < count; i++)
        {
            if (devices[i] < 0 || devices[i] >= have || std::find(v.begin(), v.end(), devices[i]) != v.end())
                return PPX_ERR_BAD_ARGUMENT;
            v.push_back(devices[i]);
        }
    }
    std::lock_guard<std::mutex> lock(g_dev_mu);
    g_devices.swap(v);
    g_devices_set = true;
    return PPX_OK;
}
catch (...)
{
    return PPX_ERR_HOST;
}

int ppx_cuda_host_get_devices(int *devices, int capacity)
try
{
    const std::vector<int> v = host_devices();
    for (int i = 0; i < (int)v.size() && i < capacity && devices; i++)
        devices[i] = v[i];
    return (int)v.size();
}
catch (...)
{
    return PPX_ERR_HOST;
}

// Measurement helper: what the host-device links of the configured devices carry when all of them copy at once.
// `host` is cut into one contiguous slice per device; every worker copies through its slice in chunks of the
// pipeline's size (device scratch of two chunks, two streams) again and again until `seconds` have passed, all
// workers starting together.  direction: 0 H2D, 1 D2H, 2 both at once (one stream each, on the two halves of the
// slice).  gbs[i] = bytes device i moved / elapsed.  No kernel runs: this is the ceiling of any host-buffer path.
int ppx_cuda_host_copy_probe(void *host, size_t bytes, int direction, double seconds, double *gbs)
try
{
    if (!host || !gbs || bytes == 0 || direction < 0 || direction > 2 || !(seconds > 0.0))
        return PPX_ERR_BAD_ARGUMENT;
    std::vector<int> devs = host_devices();
    if (devs.empty())
    {
        int cur = 0;
        const cudaError_t e = cudaGetDevice(&cur);
        if (e != cudaSuccess)
            return (int)e;
        devs.push_back(cur);
    }
    return probe_devices(devs, host, bytes, direction, seconds, gbs);
}
catch (...)
{
    return PPX_ERR_HOST;
}

// Picks, among `candidates`, the devices whose host links carry the most WHEN USED TOGETHER and makes them the device
// list of the host face.  On a multi-socket host the links are not independent: on the 8 x B200 box this was developed
// on, devices 0-3 reach host memory through a path that gives the four of them 40 GB/s of D2H in total and, while they
// use it, halves what devices 4-7 get (all eight together: 90 GB/s; devices 4-7 alone: 175 GB/s) -- so for a
// D2H-bound call four devices are faster than eight.  Greedy selection on measured rates (ppx_cuda_host_copy_probe's
// loop, `seconds` per probe): every candidate alone; then, starting from the fastest, the candidate whose addition
// raises the aggregate most is added -- taken at once, without trying the others, if it keeps at least half of its
// solo rate -- until no addition gains 3 %.  ~20 probes for 8 devices.  *aggregate_gbs = rate of the chosen set.
int ppx_cuda_host_tune_devices(const int *candidates, int count, void *host, size_t bytes, int direction, double seconds,
                               double *aggregate_gbs)
try
{
    if (!host || bytes == 0 || direction < 0 || direction > 2 || !(seconds > 0.0) || count > MAX_DEVICES)
        return PPX_ERR_BAD_ARGUMENT;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1)
        return PPX_ERR_NO_DEVICE;
    std::vector<int> cand;
    if (candidates && count > 0)
    {
        for (int i = 0; i < count; i++)
        {
            if (candidates[i] < 0 || candidates[i] >= have ||
                std::find(cand.begin(), cand.end(), candidates[i]) != cand.end())
                return PPX_ERR_BAD_ARGUMENT;
            cand.push_back(candidates[i]);
        }
    }
    else
        for (int d = 0; d < have && d < MAX_DEVICES; d++)
            cand.push_back(d);
    auto rate = [&](const std::vector<int> &set, double &agg) {
        std::vector<double> g(set.size(), 0.0);
        const int rc = probe_devices(set, host, bytes, direction, seconds, g.data());
        agg = 0.0;
        for (double x : g)
            agg += x;
        return rc;
    };
    std::vector<double> solo(cand.size(), 0.0);
    for (size_t i = 0; i < cand.size(); i++)
    {
        const int rc = rate({cand[i]}, solo[i]);
        if (rc != PPX_OK)
            return rc;
    }
    std::vector<size_t> order(cand.size());
    for (size_t i = 0; i < order.size(); i++)
        order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return solo[x] > solo[y]; });
    std::vector<int> chosen{cand[order[0]]};
    std::vector<bool> used(cand.size(), false);
    used[order[0]] = true;
    double best = solo[order[0]];
    for (;;)
    {
        size_t pick = cand.size();
        double pick_rate = best;
        for (size_t oi = 0; oi < order.size(); oi++)
        {
            const size_t i = order[oi];
            if (used[i])
                continue;
            std::vector<int> trial = chosen;
            trial.push_back(cand[i]);
            double agg = 0.0;
            const int rc = rate(trial, agg);
            if (rc != PPX_OK)
                return rc;
            if (agg > pick_rate)
            {
                pick = i;
                pick_rate = agg;
            }
            if (agg - best >= 0.5 * solo[i])
                break; // this link is (nearly) independent of the chosen ones: take it
        }
        if (pick == cand.size() || pick_rate < 1.03 * best)
            break;
        used[pick] = true;
        chosen.push_back(cand[pick]);
        best = pick_rate;
    }
    std::sort(chosen.begin(), chosen.end());
    if (aggregate_gbs)
        *aggregate_gbs = best;
    std::lock_guard<std::mutex> lock(g_dev_mu);
    g_devices = chosen;
    g_devices_set = true;
    return PPX_OK;
}
catch (...)
{
    return PPX_ERR_HOST;
}

int ppx_cuda_host_matmul(int m, int n, int l, const double *A, const double *B, double *C, size_t count)
{
    if (m < 1 || n < 1 || l < 1 || m > 8 || n > 8 || l > 8)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}, {B, (size_t)n * l * D}}, {{C, (size_t)m * l * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_matmul(m, n, l, (cd)i[0], (cd)i[1], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_so3_exp(const double *w, double *R, size_t n)
{
    return pipeline({{w, 3 * D}}, {{R, 9 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_so3_exp((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SO3_log(const double *R, double *w, size_t n)
{
    return pipeline({{R, 9 * D}}, {{w, 3 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SO3_log((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_se3_exp(const double *xi, double *T, size_t n)
{
    return pipeline({{xi, 6 * D}}, {{T, 16 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_se3_exp((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_log(const double *T, double *xi, size_t n)
{
    return pipeline({{T, 16 * D}}, {{xi, 6 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_log((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_se3_exp_log(const double *xi, double *xi_out, size_t n)
{
    return pipeline({{xi, 6 * D}}, {{xi_out, 6 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_se3_exp_log((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_mul(const double *A, const double *B, double *C, size_t n)
{
    return pipeline({{A, 16 * D}, {B, 16 * D}}, {{C, 16 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_mul((cd)i[0], (cd)i[1], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_inv(const double *T, double *Ti, size_t n)
{
    return pipeline({{T, 16 * D}}, {{Ti, 16 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_inv((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_SE3_Adt(const double *T, double *Ad, size_t n)
{
    return pipeline({{T, 16 * D}}, {{Ad, 36 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_SE3_Adt((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_se3_adt(const double *xi, double *ad, size_t n)
{
    return pipeline({{xi, 6 * D}}, {{ad, 36 * D}}, n, [](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_se3_adt((cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_so3_jac(int which, const double *w, double *J, size_t n)
{
    return pipeline({{w, 3 * D}}, {{J, 9 * D}}, n, [which](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_so3_jac(which, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_poe_fk_jacobian(const double *screws, const double *home, int njoints, const double *q,
                                  double *T, double *Js, size_t n)
{
    if (njoints < 1 || njoints > PPX_MAX_JOINTS)
        return PPX_ERR_BAD_SHAPE;
    const size_t nj = (size_t)njoints;
    return pipeline({{q, nj * D}}, {{T, 16 * D}, {Js, 6 * nj * D}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_poe_fk_jacobian(screws, home, njoints, (cd)i[0], (md)o[0], (md)o[1], c,
                                                              PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_ik_newton_step(const double *screws, const double *home, int njoints, const double *q,
                                 const double *Ttarget, double *q_out, double *Vs, size_t n)
{
    if (njoints < 3 || njoints > 6)
        return PPX_ERR_BAD_SHAPE;
    const size_t nj = (size_t)njoints;
    return pipeline({{q, nj * D}, {Ttarget, 16 * D}}, {{q_out, nj * D}, {Vs, 6 * D}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_ik_newton_step(screws, home, njoints, (cd)i[0], (cd)i[1], (md)o[0],
                                                             (md)o[1], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_ik_solve(const double *screws, const double *home, int njoints, const double *q0,
                           const double *Ttarget, int max_iter, double *q_out, int32_t *iters, int8_t *status,
                           size_t n)
{
    if (njoints != 6)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{q0, 6 * D}, {Ttarget, 16 * D}}, {{q_out, 6 * D}, {iters, sizeof(int32_t)}, {status, 1}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_ik_solve(screws, home, njoints, (cd)i[0], (cd)i[1], max_iter, (md)o[0],
                                                       (int32_t *)o[1], (int8_t *)o[2], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_linsolve_qr(int m, int n, const double *A, const double *b, double *x, int8_t *status,
                              size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}, {b, (size_t)m * D}}, {{x, (size_t)n * D}, {status, 1}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_linsolve_qr(m, n, (cd)i[0], (cd)i[1], (md)o[0], (int8_t *)o[1], c,
                                                          PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_svd(int m, int n, const double *A, double *U, double *w, double *V, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}}, {{U, (size_t)m * n * D}, {w, (size_t)n * D}, {V, (size_t)n * n * D}},
                    count, [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_svd(m, n, (cd)i[0], (md)o[0], (md)o[1], (md)o[2], c, PPX_LAYOUT_AOS, 0,
                                                  s);
                    });
}

int ppx_cuda_host_linsolve_svd(int m, int n, const double *A, const double *b, double *x, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}, {b, (size_t)m * D}}, {{x, (size_t)n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_linsolve_svd(m, n, (cd)i[0], (cd)i[1], (md)o[0], c, PPX_LAYOUT_AOS, 0,
                                                           s);
                    });
}

int ppx_cuda_host_eig_sym(int n, const double *S, double *d, double *z, int sorted, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{S, (size_t)n * n * D}}, {{d, (size_t)n * D}, {z, (size_t)n * n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_eig_sym(n, (cd)i[0], (md)o[0], (md)o[1], sorted, c, PPX_LAYOUT_AOS, 0,
                                                      s);
                    });
}

int ppx_cuda_host_linsolve_lu(int n, const double *A, const double *b, double *x, int8_t *status, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)n * n * D}, {b, (size_t)n * D}}, {{x, (size_t)n * D}, {status, 1}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_linsolve_lu(n, (cd)i[0], (cd)i[1], (md)o[0], (int8_t *)o[1], c,
                                                          PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_inverse(int n, const double *A, double *Ai, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)n * n * D}}, {{Ai, (size_t)n * n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_inverse(n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_det(int n, const double *A, double *det, size_t count)
{
    if (n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)n * n * D}}, {{det, D}}, count, [=](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_det(n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_pinv(int m, int n, const double *A, double *P, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}}, {{P, (size_t)m * n * D}}, count,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_pinv(m, n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
                    });
}

int ppx_cuda_host_norm2(int m, int n, const double *A, double *sv, size_t count)
{
    if (m < 1 || n < 1)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{A, (size_t)m * n * D}}, {{sv, D}}, count, [=](size_t c, void **i, void **o, void *s) {
        return ppx_cuda_batch_norm2(m, n, (cd)i[0], (md)o[0], c, PPX_LAYOUT_AOS, 0, s);
    });
}

int ppx_cuda_host_ik_codo(const double *screws, const double *home, int njoints, const double *lo, const double *hi,
                          const double *Ttarget, const double *init, double *q_out, double *x_raw, double *fnorm,
                          int8_t *status, size_t n)
{
    if (njoints != 6)
        return PPX_ERR_BAD_SHAPE;
    return pipeline({{Ttarget, 16 * D}, {init, 6 * D}}, {{q_out, 6 * D}, {x_raw, 6 * D}, {fnorm, D}, {status, 1}}, n,
                    [=](size_t c, void **i, void **o, void *s) {
                        return ppx_cuda_batch_ik_codo(screws, home, njoints, lo, hi, (cd)i[0], (cd)i[1], (md)o[0], (md)o[1],
                                                      (md)o[2], (int8_t *)o[3], c, PPX_LAYOUT_AOS, 0, s);
                    },
                    16);
}

} // extern "C"
