// host_stage.inl -- pageable host arrays behind the host-pointer calls (included by pixsht.cu).
//
// The reference hands libsharp2 ordinary Julia arrays (src/transforms.jl:101-106, 185-194): pageable memory.  A cudaMemcpyAsync
// from / to pageable memory is staged by the driver on the calling thread at a few GB/s and blocks it, which serialises the
// three-stream pipeline of execute_host (measured: C4 1704 ms against 588 ms from page-locked buffers, profiles/r02).  Callers
// who can should allocate through pixsht_host_alloc or page-lock their arrays with pixsht_host_register; for everybody else
// the library stages pageable arrays itself:
//
//   H2D:  copy threads fill a ring of page-locked slots from the caller's array (a slot is split over the worker pool), the
//         DMA of a slot is enqueued behind a host function that waits for its fill, a second host function frees the slot;
//   D2H:  the DMA into a slot is enqueued behind a host function that waits until the slot's previous contents have been
//         drained, a second host function hands the slot to the copy threads, which drain it into the caller's array.
//
// Everything is enqueued by the calling thread in the same order as the plain cudaMemcpyAsync calls it replaces, so events
// recorded after a staged copy mean what they meant before.  The host functions wait only for CPU work of the copy threads
// (never for CUDA work), and fill / drain have separate dispatcher threads, so there is no cycle.
#pragma once
#ifndef PIXSHT_EMU
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>

namespace pixsht_stage {

// ---- process-wide pool that splits one memcpy over several threads ------------------------------------------------------
class CopyPool {
public:
    static CopyPool& get() { static CopyPool p; return p; }
    int width() const { return (int)th_.size(); }
    // blocking: dst <- src, split over the pool (callers are the dispatcher threads of the stagers)
    void copy(void* dst, const void* src, size_t n)
    {
        const int W = width();
        if (W <= 1 || n < (1u << 20)) { memcpy(dst, src, n); return; }
        auto done = std::make_shared<std::atomic<int>>(0);
        const size_t per = ((n + W - 1) / W + 4095) & ~(size_t)4095;
        int parts = 0;
        {
            std::lock_guard<std::mutex> g(mu_);
            for (size_t off = 0; off < n; off += per) {
                const size_t len = std::min(per, n - off);
                q_.push_back([=]() { memcpy((char*)dst + off, (const char*)src + off, len); done->fetch_add(1, std::memory_order_release); });
                ++parts;
            }
        }
        cv_.notify_all();
        // help out instead of sleeping: the dispatcher takes tasks too
        for (;;) {
            std::function<void()> f;
            {
                std::lock_guard<std::mutex> g(mu_);
                if (!q_.empty()) { f = std::move(q_.front()); q_.pop_front(); }
            }
            if (f) { f(); continue; }
            if (done->load(std::memory_order_acquire) >= parts) break;
            std::this_thread::yield();
        }
    }

private:
    CopyPool()
    {
        unsigned hw = std::thread::hardware_concurrency();
        int n = env_int("PIXSHT_STAGE_THREADS", hw >= 32 ? 12 : (hw >= 16 ? 8 : (hw >= 8 ? 4 : 2)));
        n = std::max(1, std::min(n, 64));
        for (int i = 0; i < n; ++i)
            th_.emplace_back([this]() {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> g(mu_);
                        cv_.wait(g, [this]() { return stop_ || !q_.empty(); });
                        if (stop_ && q_.empty()) return;
                        f = std::move(q_.front()); q_.pop_front();
                    }
                    f();
                }
            });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) if (t.joinable()) t.join();
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> th_;
    bool stop_ = false;
};

// ---- per-plan ring of page-locked slots, one ring and one dispatcher thread per direction --------------------------------
class Stager {
public:
    Stager()
    {
        chunk_ = (size_t)std::max(1, std::min(256, env_int("PIXSHT_STAGE_CHUNK_MB", 64))) << 20;
        nslot_ = std::max(2, std::min(8, env_int("PIXSHT_STAGE_SLOTS", 4)));
    }
    ~Stager()
    {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : disp_) if (t.joinable()) t.join();
        for (int d = 0; d < 2; ++d) for (void* p : slot_[d]) if (p) cudaFreeHost(p);
        (void)cudaGetLastError();
    }
    size_t chunk() const { return chunk_; }

    // H2D of a pageable array: enqueued on `s` like one cudaMemcpyAsync
    cudaError_t h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s)
    {
        cudaError_t e = ensure(0); if (e != cudaSuccess) return e;
        for (size_t off = 0; off < bytes; off += chunk_) {
            const size_t n = std::min(chunk_, bytes - off);
            const unsigned long long q = seq_[0]++;
            const int slot = (int)(q % nslot_);
            const unsigned long long gen = q / nslot_;
            push(0, [this, slot, gen, src_host, off, n]() {
                wait_ge(freed_[slot], gen);                                   // the slot's previous DMA has read it
                CopyPool::get().copy(slot_[0][slot], (const char*)src_host + off, n);
                set(filled_[slot], gen + 1);
            });
            e = cudaLaunchHostFunc(s, &Stager::cb, note(Note{this, 0, slot, gen, nullptr, 0})); if (e != cudaSuccess) return e;      // wait for the fill
            e = cudaMemcpyAsync((char*)dst_dev + off, slot_[0][slot], n, cudaMemcpyHostToDevice, s); if (e != cudaSuccess) return e;
            e = cudaLaunchHostFunc(s, &Stager::cb, note(Note{this, 1, slot, gen, nullptr, 0})); if (e != cudaSuccess) return e;      // free the slot
        }
        return cudaSuccess;
    }
    // D2H into a pageable array; the caller's memory is complete after the stream has been synchronised AND drain_wait()
    cudaError_t d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s)
    {
        cudaError_t e = ensure(1); if (e != cudaSuccess) return e;
        for (size_t off = 0; off < bytes; off += chunk_) {
            const size_t n = std::min(chunk_, bytes - off);
            const unsigned long long q = seq_[1]++;
            const int slot = (int)(q % nslot_);
            const unsigned long long gen = q / nslot_;
            e = cudaLaunchHostFunc(s, &Stager::cb, note(Note{this, 2, slot, gen, nullptr, 0})); if (e != cudaSuccess) return e;      // previous contents drained
            e = cudaMemcpyAsync(slot_[1][slot], (const char*)src_dev + off, n, cudaMemcpyDeviceToHost, s); if (e != cudaSuccess) return e;
            { std::lock_guard<std::mutex> g(mu_); ++drain_pending_; }
            e = cudaLaunchHostFunc(s, &Stager::cb, note(Note{this, 3, slot, gen, (char*)dst_host + off, n}));                        // hand the slot to the drain thread
            if (e != cudaSuccess) { std::lock_guard<std::mutex> g(mu_); --drain_pending_; return e; }
        }
        return cudaSuccess;
    }
    // after the streams have been synchronised: every drained byte is in the caller's array; the call's notes can go
    void drain_wait()
    {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this]() { return drain_pending_ == 0; });
        notes_.clear();
    }

private:
    struct Note { Stager* st; int kind; int slot; unsigned long long gen; void* dst; size_t n; };
    static void CUDART_CB cb(void* p)
    {
        Note* x = static_cast<Note*>(p);
        Stager* S = x->st;
        switch (x->kind) {
        case 0: S->wait_ge(S->filled_[x->slot], x->gen + 1); break;
        case 1: S->set(S->freed_[x->slot], x->gen + 1); break;
        case 2: S->wait_ge(S->drained_[x->slot], x->gen); break;
        default: {
            const int slot = x->slot; const unsigned long long gen = x->gen; void* dst = x->dst; const size_t n = x->n;
            S->push(1, [S, slot, gen, dst, n]() {
                CopyPool::get().copy(dst, S->slot_[1][slot], n);
                std::lock_guard<std::mutex> g(S->mu_);
                S->drained_[slot] = gen + 1; --S->drain_pending_;
                S->cv_.notify_all();
            });
        }
        }
    }
    void* note(const Note& n) { std::lock_guard<std::mutex> g(mu_); notes_.push_back(n); return &notes_.back(); }   // deque: stable addresses
    void wait_ge(unsigned long long& v, unsigned long long want)
    {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&]() { return v >= want || stop_; });
    }
    void set(unsigned long long& v, unsigned long long val)
    {
        { std::lock_guard<std::mutex> g(mu_); v = val; }
        cv_.notify_all();
    }
    void push(int dir, std::function<void()> f)
    {
        { std::lock_guard<std::mutex> g(mu_); jobs_[dir].push_back(std::move(f)); }
        cv_.notify_all();
    }
    cudaError_t ensure(int dir)
    {
        if (!slot_[dir].empty()) return cudaSuccess;
        slot_[dir].assign(nslot_, nullptr);
        for (int i = 0; i < nslot_; ++i) {
            const cudaError_t e = cudaHostAlloc(&slot_[dir][i], chunk_, cudaHostAllocPortable);
            if (e != cudaSuccess) { slot_[dir].clear(); return e; }
        }
        if (dir == 0) { filled_.assign(nslot_, 0); freed_.assign(nslot_, 0); } else drained_.assign(nslot_, 0);
        disp_.emplace_back([this, dir]() {
            for (;;) {
                std::function<void()> f;
                {
                    std::unique_lock<std::mutex> g(mu_);
                    cv_.wait(g, [&]() { return stop_ || !jobs_[dir].empty(); });
                    if (stop_ && jobs_[dir].empty()) return;
                    f = std::move(jobs_[dir].front()); jobs_[dir].pop_front();
                }
                f();
            }
        });
        return cudaSuccess;
    }
    size_t chunk_ = 64u << 20;
    int nslot_ = 4;
    std::vector<void*> slot_[2];
    std::vector<unsigned long long> filled_, freed_, drained_;
    unsigned long long seq_[2] = {0, 0};
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> jobs_[2];
    std::deque<Note> notes_;
    std::vector<std::thread> disp_;
    long long drain_pending_ = 0;
    bool stop_ = false;
};

// pageable = ordinary host memory the CUDA driver does not know (not page-locked, not device, not managed)
static bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace pixsht_stage
#endif
