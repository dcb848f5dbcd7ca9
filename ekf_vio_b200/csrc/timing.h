// Per-kernel CUDA-event timing used by bench.py's roofline leg: events are recorded on the
// launching stream around each kernel and resolved after a synchronize.
#pragma once
#include <cuda_runtime.h>

#include <vector>

struct KernelTimer {
    static constexpr int SLOTS = 8;
    bool on = false;
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    double ms[SLOTS] = {0};
    long long cnt[SLOTS] = {0};
    cudaEvent_t cur_a = nullptr;
    int cur_id = -1;

    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void begin(int id, cudaStream_t st) {
        if (!on) return;
        // timing events must not become nodes of a caller's CUDA graph (ekfvio_vio_add_frame records these entry points)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cur_id = -1; return; }
        cur_a = get(); cur_id = id;
        cudaEventRecord(cur_a, st);
    }
    void end(cudaStream_t st) {
        if (!on || cur_id < 0) return;
        cudaEvent_t b = get();
        cudaEventRecord(b, st);
        pending.push_back({cur_id, cur_a, b});
        cur_id = -1;
    }
    void resolve() {
        for (auto& r : pending) {
            cudaEventSynchronize(r.b);
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.id >= 0 && r.id < SLOTS) { ms[r.id] += t; cnt[r.id] += 1; }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        pending.clear();
    }
    void reset() { resolve(); for (int i = 0; i < SLOTS; ++i) { ms[i] = 0; cnt[i] = 0; } }
    ~KernelTimer() { resolve(); for (auto e : pool) cudaEventDestroy(e); }
};
