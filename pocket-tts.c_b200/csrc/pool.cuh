// pool.cuh - host-side multi-GPU driver of the engine: one engine + one host thread per device, the
// utterance list sorted by length, cut into calls and dealt to the devices, results gathered in the original order. No collective,
// no peer traffic: utterances are independent and the weights are replicated (~200 MB per device).
// The reference has no counterpart (ptts_generate is single-utterance, ptts.h:75-76; ptts_cuda.c drives one device).
#pragma once
#include <algorithm>
#include <numeric>
#include <string>
#include <thread>

namespace ptts {

struct PoolCall {
    std::vector<int> idx;      // original utterance indices, longest first
    int cap = 0;               // frames of the longest utterance of the call
    int device = 0;            // slot in the pool
    bool direct = false;       // utterances are a contiguous ascending index range and cap == the job's cap: write in place
};

}  // namespace ptts

struct ptts_b200_pool {
    std::vector<ptts_b200_engine *> eng;
    std::vector<int> dev;
    // per device: two pinned PCM staging buffers (grow-only, kept for the life of the pool). A call whose utterances are not one
    // contiguous index range of full length lands in a staging buffer; its rows are copied to their final places by a helper
    // thread while the device already runs the next call into the other buffer.
    struct Stage { float *buf[2] = {nullptr, nullptr}; size_t cap[2] = {0, 0}; };
    std::vector<Stage> stage;
};

extern "C" int ptts_b200_pool_create(ptts_b200_pool **out, const int *devices, int n_devices, const ptts_b200_tensor *tensors, int n_tensors) {
    using namespace ptts;
    PTTS_CHECK(out && tensors && n_tensors > 0, "pool_create: bad arguments");
    *out = nullptr;
    std::vector<int> dev;
    if (devices && n_devices > 0) dev.assign(devices, devices + n_devices);
    else {
        int nd = 0;
        cudaError_t r = cudaGetDeviceCount(&nd);
        PTTS_CHECK(r == cudaSuccess && nd > 0, "no CUDA device available (%s): this engine has no CPU fallback", cudaGetErrorString(r));
        dev.resize(nd);
        std::iota(dev.begin(), dev.end(), 0);
    }
    ptts_b200_pool *p = new ptts_b200_pool();
    p->dev = dev;
    p->eng.assign(dev.size(), nullptr);
    p->stage.assign(dev.size(), ptts_b200_pool::Stage());
    std::vector<std::string> err(dev.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < dev.size(); i++)   // weight conversion and upload of the replicas run side by side
        th.emplace_back([&, i] {
            if (ptts_b200_create(&p->eng[i], dev[i], tensors, n_tensors) != 0) err[i] = ptts_b200_last_error();
        });
    for (auto &t : th) t.join();
    for (size_t i = 0; i < dev.size(); i++)
        if (!p->eng[i]) {
            set_error("pool_create: device %d: %s", dev[i], err[i].c_str());
            for (auto *e : p->eng) ptts_b200_destroy(e);
            delete p;
            return -1;
        }
    *out = p;
    return 0;
}

extern "C" void ptts_b200_pool_destroy(ptts_b200_pool *p) {
    if (!p) return;
    for (size_t d = 0; d < p->eng.size(); d++) {
        if (p->eng[d]) cudaSetDevice(p->eng[d]->device);
        for (int k = 0; k < 2; k++)
            if (p->stage[d].buf[k]) cudaFreeHost(p->stage[d].buf[k]);
    }
    for (auto *e : p->eng) ptts_b200_destroy(e);
    delete p;
}
extern "C" int ptts_b200_pool_size(const ptts_b200_pool *p) { return p ? (int)p->eng.size() : 0; }
extern "C" ptts_b200_engine *ptts_b200_pool_engine(ptts_b200_pool *p, int i) { return p && i >= 0 && i < (int)p->eng.size() ? p->eng[i] : nullptr; }

extern "C" int ptts_b200_pool_generate(ptts_b200_pool *p, int n, const int *tokens_flat, const int *token_offsets, const float *cond_prefix,
                                       int cond_len, const int *max_frames, const int64_t *seeds, const int *eos_after,
                                       const ptts_b200_gen_params *params, int max_batch, float *out_latents, int *out_frames_used,
                                       float *out_audio) {
    using namespace ptts;
    PTTS_CHECK(p && !p->eng.empty() && n >= 1 && tokens_flat && token_offsets && max_frames && seeds && params, "pool_generate: bad arguments");
    const int nd = (int)p->eng.size();
    if (max_batch <= 0 || max_batch > 1024) max_batch = 1024;
    int F = 0;
    for (int i = 0; i < n; i++) { PTTS_CHECK(max_frames[i] >= 1, "pool_generate: utterance %d has max_frames < 1", i); F = std::max(F, max_frames[i]); }
    // ---- bucket by length (ptts_estimate_frames, ptts.c:285-291, is what callers put into max_frames): longest first, so every call
    // runs the lock-step loop and the Mimi decoder only to its own longest utterance
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return max_frames[a] > max_frames[b]; });
    // calls: at most max_batch utterances, and at least one call per device when there is enough work to share
    int per = std::min(max_batch, std::max(16, (n + nd - 1) / nd));
    per = std::min(max_batch, (per + 15) / 16 * 16);
    std::vector<PoolCall> calls;
    for (int i = 0; i < n; i += per) {
        PoolCall c;
        c.idx.assign(order.begin() + i, order.begin() + std::min(n, i + per));
        c.cap = max_frames[c.idx[0]];
        bool contiguous = true;
        for (size_t k = 1; k < c.idx.size(); k++) contiguous = contiguous && c.idx[k] == c.idx[0] + (int)k;
        c.direct = contiguous && c.cap == F;
        calls.push_back(std::move(c));
    }
    // longest-processing-time assignment: cost ~ frames x (fixed per-step latency + per-utterance share)
    std::vector<double> load(nd, 0.0);
    std::vector<int> by_cost(calls.size());
    std::iota(by_cost.begin(), by_cost.end(), 0);
    auto cost = [&](const PoolCall &c) { return (double)c.cap * (256.0 + (double)c.idx.size()); };
    std::stable_sort(by_cost.begin(), by_cost.end(), [&](int a, int b) { return cost(calls[a]) > cost(calls[b]); });
    for (int ci : by_cost) {
        const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        calls[ci].device = d;
        load[d] += cost(calls[ci]);
    }
    // ---- one host thread per device
    std::vector<std::string> err(nd);
    auto worker = [&](int d) {
        ptts_b200_engine *e = p->eng[d];
        ptts_b200_pool::Stage &sg = p->stage[d];
        std::vector<int> tok, offs, mf, ea;
        std::vector<int64_t> sd;
        // per staging buffer: the latents / frame counts of the call it holds and the thread gathering them
        std::vector<float> lat[2];
        std::vector<int> used[2];
        std::thread gather[2];
        int which = 0;
        for (const PoolCall &c : calls) {
            if (c.device != d || !err[d].empty()) continue;
            const int m = (int)c.idx.size(), w = which;
            if (gather[w].joinable()) gather[w].join();   // the previous user of this buffer has been copied out
            tok.clear(); offs.assign(1, 0); mf.resize(m); sd.resize(m); ea.resize(m); used[w].assign(m, 0);
            for (int k = 0; k < m; k++) {
                const int i = c.idx[k];
                tok.insert(tok.end(), tokens_flat + token_offsets[i], tokens_flat + token_offsets[i + 1]);
                offs.push_back((int)tok.size());
                mf[k] = max_frames[i]; sd[k] = seeds[i]; ea[k] = eos_after ? eos_after[i] : params->eos_after;
            }
            float *lat_dst = nullptr, *pcm_dst = nullptr;
            if (c.direct) {
                lat_dst = out_latents ? out_latents + (size_t)c.idx[0] * F * FL_LAT : nullptr;
                pcm_dst = out_audio ? out_audio + (size_t)c.idx[0] * F * 1920 : nullptr;
            } else {
                if (out_latents) { lat[w].resize((size_t)m * c.cap * FL_LAT); lat_dst = lat[w].data(); }
                if (out_audio) {
                    const size_t need = (size_t)m * c.cap * 1920;
                    if (need > sg.cap[w]) {
                        cudaSetDevice(e->device);
                        if (sg.buf[w]) cudaFreeHost(sg.buf[w]);
                        sg.buf[w] = nullptr; sg.cap[w] = 0;
                        if (cudaMallocHost(&sg.buf[w], need * sizeof(float)) != cudaSuccess) { err[d] = "pool_generate: pinned staging allocation failed"; break; }
                        sg.cap[w] = need;
                    }
                    pcm_dst = sg.buf[w];
                }
            }
            if (ptts_b200_generate_batch(e, m, tok.data(), offs.data(), cond_prefix, cond_len, mf.data(), sd.data(), ea.data(), params, lat_dst,
                                         used[w].data(), pcm_dst) != 0) {
                err[d] = ptts_b200_last_error();
                break;
            }
            auto scatter = [&, w, lat_dst, pcm_dst](const PoolCall *cc) {   // results to their places in the original order
                const int mm = (int)cc->idx.size();
                for (int k = 0; k < mm; k++) {
                    const int i = cc->idx[k];
                    if (out_frames_used) out_frames_used[i] = used[w][k];
                    if (cc->direct) continue;
                    if (out_latents) {
                        memcpy(out_latents + (size_t)i * F * FL_LAT, lat_dst + (size_t)k * cc->cap * FL_LAT, (size_t)cc->cap * FL_LAT * sizeof(float));
                        memset(out_latents + ((size_t)i * F + cc->cap) * FL_LAT, 0, (size_t)(F - cc->cap) * FL_LAT * sizeof(float));
                    }
                    if (out_audio) {
                        memcpy(out_audio + (size_t)i * F * 1920, pcm_dst + (size_t)k * cc->cap * 1920, (size_t)cc->cap * 1920 * sizeof(float));
                        memset(out_audio + ((size_t)i * F + cc->cap) * 1920, 0, (size_t)(F - cc->cap) * 1920 * sizeof(float));
                    }
                }
            };
            if (c.direct) scatter(&c);
            else { gather[w] = std::thread(scatter, &c); which ^= 1; }
        }
        for (int k = 0; k < 2; k++)
            if (gather[k].joinable()) gather[k].join();
    };
    std::vector<std::thread> th;
    for (int d = 1; d < nd; d++) th.emplace_back(worker, d);
    worker(0);
    for (auto &t : th) t.join();
    for (int d = 0; d < nd; d++)
        if (!err[d].empty()) { set_error("pool_generate: device %d: %s", p->dev[d], err[d].c_str()); return -1; }
    return 0;
}
