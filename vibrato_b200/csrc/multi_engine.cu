// One tokenizer over several GPUs of one node (the reference has no counterpart — a vibrato
// Worker is a single CPU thread, worker.rs:13-31 — so the contract is BASELINE.json's north_star: "split the
// batch with one NCCL broadcast of the dictionary and a gather of token spans over NVLink").
//
//   * The dictionary image is uploaded ONCE (to the first device) and broadcast to the others with
//     ncclBroadcast; NCCL is bound at run time (dlopen of libnccl.so.2) so that the single-GPU library keeps no
//     link-time dependency on it.  Where NCCL cannot be loaded the same bytes travel by cudaMemcpyPeerAsync
//     (still NVLink / NVSwitch, still one PCIe upload).
//   * A batch is cut into contiguous shards of about equal BYTES (sentences are independent units; work is
//     proportional to characters), one per device; every device runs the ordinary single-device engine on its
//     shard from its own host thread (pinned to the CPUs of the GPU's NUMA node), so H2D copies, kernels and
//     D2H copies of different devices overlap and use their own PCIe links.
//   * Host results: once every shard's token count is known the shards' token records are copied straight into
//     ONE pinned result at their final positions (no host-side merge pass); token offsets are made global on
//     the device before they leave it.  The result is indistinguishable from a single-device one.
//   * Device-resident batches (vbt_tokenize_batch_device): the input sits on the first device, shards reach the
//     other devices by peer copies, and the token records come back to the first device with grouped
//     ncclSend / ncclRecv (the "gather of token spans over NVLink"; peer copies without NCCL).
#include "engine.hpp"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and enums only: every entry point is resolved with dlsym
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

namespace vbt {

namespace {

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(kCuda, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(x) cuda_check((x), #x)

// ---- NCCL, bound at run time ---------------------------------------------------------------------
struct Nccl {
    void* so = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    int version = 0;

    static Nccl* get() {  // nullptr when the library (or one of its symbols) is missing
        static Nccl inst;
        static bool tried = false;
        if (!tried) {
            tried = true;
            const char* off = std::getenv("VBT_NO_NCCL");
            if (off && off[0] == '1') return nullptr;
            inst.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
            if (inst.so) {
                bool ok = true;
                auto sym = [&](const char* n) {
                    void* p = dlsym(inst.so, n);
                    ok = ok && p;
                    return p;
                };
                inst.GetVersion = reinterpret_cast<decltype(inst.GetVersion)>(sym("ncclGetVersion"));
                inst.GetErrorString = reinterpret_cast<decltype(inst.GetErrorString)>(sym("ncclGetErrorString"));
                inst.CommInitAll = reinterpret_cast<decltype(inst.CommInitAll)>(sym("ncclCommInitAll"));
                inst.CommDestroy = reinterpret_cast<decltype(inst.CommDestroy)>(sym("ncclCommDestroy"));
                inst.Broadcast = reinterpret_cast<decltype(inst.Broadcast)>(sym("ncclBroadcast"));
                inst.Send = reinterpret_cast<decltype(inst.Send)>(sym("ncclSend"));
                inst.Recv = reinterpret_cast<decltype(inst.Recv)>(sym("ncclRecv"));
                inst.GroupStart = reinterpret_cast<decltype(inst.GroupStart)>(sym("ncclGroupStart"));
                inst.GroupEnd = reinterpret_cast<decltype(inst.GroupEnd)>(sym("ncclGroupEnd"));
                if (!ok) {
                    dlclose(inst.so);
                    inst.so = nullptr;
                } else {
                    inst.GetVersion(&inst.version);
                }
            }
        }
        return inst.so ? &inst : nullptr;
    }
    void check(ncclResult_t r, const char* what) const {
        if (r != ncclSuccess) throw Error(kCuda, std::string(what) + ": " + GetErrorString(r));
    }
};
#define NK(x) nccl_->check((x), #x)

// ---- CPU affinity: the cores of the NUMA node a GPU hangs off ---------------------------------------
std::vector<int> parse_cpulist(const std::string& s) {
    std::vector<int> out;
    std::stringstream ss(s);
    std::string part;
    while (std::getline(ss, part, ',')) {
        int a = 0, b = 0;
        if (std::sscanf(part.c_str(), "%d-%d", &a, &b) == 2) {
            for (int c = a; c <= b; ++c) out.push_back(c);
        } else if (std::sscanf(part.c_str(), "%d", &a) == 1) {
            out.push_back(a);
        }
    }
    return out;
}

std::vector<int> numa_cpus_of_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return {};
    for (char* p = bus; *p; ++p) *p = char(std::tolower(*p));
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
    int node = -1;
    if (!(f >> node) || node < 0) return {};
    std::ifstream g("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!std::getline(g, list)) return {};
    return parse_cpulist(list);
}

void pin_this_thread(const std::vector<int>& cpus) {
    if (cpus.empty()) return;
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
    int n = 0;
    for (int c : cpus)
        if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) {
            CPU_SET(c, &want);
            ++n;
        }
    if (n) pthread_setaffinity_np(pthread_self(), sizeof want, &want);  // never widens the process's own mask
}

class MultiEngine final : public Engine {
   public:
    MultiEngine(const std::vector<int>& devices, const uint8_t* host_blob, uint64_t n_bytes, bool ignore_space,
                uint64_t max_grouping_len)
        : dev_(devices) {
        const int n = int(dev_.size());
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            throw Error(kNoDevice, "no CUDA device available: the tokenizer has no CPU fallback");
        for (int i = 0; i < n; ++i) {
            if (dev_[i] < 0 || dev_[i] >= count) throw Error(kInvalidArgument, "device ordinal out of range");
            for (int j = 0; j < i; ++j)
                if (dev_[j] == dev_[i]) throw Error(kInvalidArgument, "the same device is listed twice");
        }
        nccl_ = Nccl::get();
        stream_.resize(n);
        image_.resize(n, nullptr);
        cpus_.resize(n);
        for (int i = 0; i < n; ++i) {
            CK(cudaSetDevice(dev_[i]));
            CK(cudaStreamCreateWithFlags(&stream_[i], cudaStreamNonBlocking));
            CK(cudaMalloc(&image_[i], n_bytes));
            cpus_[i] = numa_cpus_of_device(dev_[i]);
            for (int j = 0; j < n; ++j) {  // peer access for the device-resident route (NVLink / NVSwitch)
                if (i == j) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, dev_[i], dev_[j]);
                if (can) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(dev_[j], 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                }
            }
        }
        // one PCIe upload, then the image goes from device to device
        CK(cudaSetDevice(dev_[0]));
        CK(cudaMemcpyAsync(image_[0], host_blob, n_bytes, cudaMemcpyHostToDevice, stream_[0]));
        CK(cudaStreamSynchronize(stream_[0]));
        if (n > 1) {
            if (nccl_) {
                comm_.resize(n);
                NK(nccl_->CommInitAll(comm_.data(), n, dev_.data()));
                NK(nccl_->GroupStart());
                for (int i = 0; i < n; ++i) {
                    CK(cudaSetDevice(dev_[i]));
                    NK(nccl_->Broadcast(image_[i], image_[i], n_bytes, ncclUint8, 0, comm_[i], stream_[i]));
                }
                NK(nccl_->GroupEnd());
                transport_ = "nccl " + std::to_string(nccl_->version);
            } else {
                for (int i = 1; i < n; ++i)
                    CK(cudaMemcpyPeerAsync(image_[i], dev_[i], image_[0], dev_[0], n_bytes, stream_[i]));
                transport_ = "cudaMemcpyPeer";
            }
            for (int i = 0; i < n; ++i) {
                CK(cudaSetDevice(dev_[i]));
                CK(cudaStreamSynchronize(stream_[i]));
            }
        } else {
            transport_ = "single device";
        }
        for (int i = 0; i < n; ++i) {
            eng_.push_back(Engine::create(dev_[i], nullptr, reinterpret_cast<uint64_t>(image_[i]), n_bytes, ignore_space,
                                          max_grouping_len));
            if (n > 1) eng_.back()->set_option("chunk_sentences", 0);  // a shard runs as one piece: overlap comes from the devices
        }
        std::memset(stage_ms_, 0, sizeof stage_ms_);
        std::memset(counters_, 0, sizeof counters_);
    }

    ~MultiEngine() override {
        eng_.clear();
        for (size_t i = 0; i < dev_.size(); ++i) {
            cudaSetDevice(dev_[i]);
            if (nccl_ && i < comm_.size() && comm_[i]) nccl_->CommDestroy(comm_[i]);
            cudaFree(image_[i]);
            cudaStreamDestroy(stream_[i]);
        }
        for (auto* r : pool_) {
            pinned_free(r->tok_off);
            pinned_free(r->tokens);
            delete r;
        }
        pinned_free(h_off_);
        cudaSetDevice(dev_[0]);
        if (gather_tokens_) cudaFree(gather_tokens_);
        if (gather_off_) cudaFree(gather_off_);
    }

    // ---- host batches --------------------------------------------------------------------------------
    HostResult* run_host(const char* utf8, const uint64_t* byte_off, uint64_t n_sent) override {
        const int n = int(dev_.size());
        if (n == 1) {  // nothing to split: the device's own chunked, overlapped host pipeline
            HostResult* r = eng_[0]->run_host(utf8, byte_off, n_sent);
            collect_stats();
            return r;
        }
        const std::vector<uint64_t> cut = split_by_bytes(byte_off, n_sent, n);
        std::vector<uint64_t> n_tok(n, 0);
        run_on_all([&](int i) { n_tok[i] = eng_[i]->run_shard(utf8, byte_off + cut[i], cut[i + 1] - cut[i], -1); });
        std::vector<uint64_t> base(n + 1, 0);
        for (int i = 0; i < n; ++i) base[i + 1] = base[i] + n_tok[i];
        HostResult* r = acquire(n_sent, base[n]);
        const uint64_t tb = eng_[0]->token_bytes();
        r->token_bytes = uint32_t(tb);
        int last = n - 1;  // the shard that also delivers the closing offset: the last one, even when it is empty
        run_on_all([&](int i) {
            eng_[i]->fetch_shard(r->tok_off + cut[i], static_cast<uint8_t*>(r->tokens) + base[i] * tb, base[i], i == last);
        });
        collect_stats();
        return r;
    }
    void release(HostResult* r) override {
        if (!r) return;
        if (dev_.size() == 1)
            eng_[0]->release(r);
        else
            free_.push_back(r);
    }

    // ---- device-resident batches: input on the first device, results gathered back to it ------------------
    void run_device(uint64_t d_utf8, uint64_t d_byte_off, uint64_t n_sent, uint64_t n_bytes, uint64_t* d_tok_off,
                    uint64_t* d_tokens, uint64_t* n_tokens) override {
        const int n = int(dev_.size());
        if (n == 1) {
            eng_[0]->run_device(d_utf8, d_byte_off, n_sent, n_bytes, d_tok_off, d_tokens, n_tokens);
            collect_stats();
            return;
        }
        CK(cudaSetDevice(dev_[0]));
        // the split needs the offsets on the host (pinned staging buffer, one copy); their validity is checked on
        // the devices like for any device-resident batch (k_count_chars)
        if ((n_sent + 1) * 8 > h_off_cap_) {
            pinned_free(h_off_);
            h_off_cap_ = size_t(double((n_sent + 1) * 8) * 1.25) + 64;
            h_off_ = static_cast<uint64_t*>(pinned_alloc(h_off_cap_));
        }
        CK(cudaMemcpyAsync(h_off_, reinterpret_cast<const void*>(d_byte_off), (n_sent + 1) * 8, cudaMemcpyDeviceToHost, stream_[0]));
        CK(cudaStreamSynchronize(stream_[0]));
        const std::vector<uint64_t> cut = split_by_bytes(h_off_, n_sent, n);
        std::vector<uint64_t> n_tok(n, 0);
        run_on_all([&](int i) {
            n_tok[i] = eng_[i]->run_shard(reinterpret_cast<const char*>(d_utf8), h_off_ + cut[i], cut[i + 1] - cut[i],
                                          dev_[0]);
        });
        std::vector<uint64_t> base(n + 1, 0);
        for (int i = 0; i < n; ++i) base[i + 1] = base[i] + n_tok[i];
        run_on_all([&](int i) { eng_[i]->rebase_shard(base[i]); });
        // gather on the first device: token records and offsets of every shard at their final positions
        CK(cudaSetDevice(dev_[0]));
        if (base[n] * 24 + 24 > gather_tok_cap_) {
            if (gather_tokens_) CK(cudaFree(gather_tokens_));
            gather_tok_cap_ = size_t(double(base[n] * 24) * 1.25) + 4096;
            CK(cudaMalloc(&gather_tokens_, gather_tok_cap_));
        }
        if ((n_sent + 1) * 8 > gather_off_cap_) {
            if (gather_off_) CK(cudaFree(gather_off_));
            gather_off_cap_ = size_t(double((n_sent + 1) * 8) * 1.25) + 4096;
            CK(cudaMalloc(&gather_off_, gather_off_cap_));
        }
        std::vector<uint64_t> src_off(n), src_tok(n);
        for (int i = 0; i < n; ++i) eng_[i]->shard_outputs(&src_off[i], &src_tok[i]);
        const uint64_t tb = eng_[0]->token_bytes();
        auto dst_tok = [&](int i) { return static_cast<uint8_t*>(gather_tokens_) + base[i] * tb; };
        auto dst_off = [&](int i) { return static_cast<uint8_t*>(gather_off_) + cut[i] * 8; };
        auto off_bytes = [&](int i) { return (cut[i + 1] - cut[i] + (i == n - 1 ? 1 : 0)) * 8; };
        // shard 0 is already on the first device
        CK(cudaMemcpyAsync(dst_tok(0), reinterpret_cast<const void*>(src_tok[0]), n_tok[0] * tb, cudaMemcpyDeviceToDevice, stream_[0]));
        CK(cudaMemcpyAsync(dst_off(0), reinterpret_cast<const void*>(src_off[0]), off_bytes(0), cudaMemcpyDeviceToDevice, stream_[0]));
        if (n > 1 && nccl_) {
            NK(nccl_->GroupStart());
            for (int i = 1; i < n; ++i) {
                CK(cudaSetDevice(dev_[i]));
                if (n_tok[i]) NK(nccl_->Send(reinterpret_cast<const void*>(src_tok[i]), n_tok[i] * tb, ncclUint8, 0, comm_[i], stream_[i]));
                NK(nccl_->Send(reinterpret_cast<const void*>(src_off[i]), off_bytes(i), ncclUint8, 0, comm_[i], stream_[i]));
                CK(cudaSetDevice(dev_[0]));
                if (n_tok[i]) NK(nccl_->Recv(dst_tok(i), n_tok[i] * tb, ncclUint8, i, comm_[0], stream_[0]));
                NK(nccl_->Recv(dst_off(i), off_bytes(i), ncclUint8, i, comm_[0], stream_[0]));
            }
            NK(nccl_->GroupEnd());
        } else {
            for (int i = 1; i < n; ++i) {
                if (n_tok[i]) CK(cudaMemcpyPeerAsync(dst_tok(i), dev_[0], reinterpret_cast<const void*>(src_tok[i]), dev_[i], n_tok[i] * tb, stream_[0]));
                CK(cudaMemcpyPeerAsync(dst_off(i), dev_[0], reinterpret_cast<const void*>(src_off[i]), dev_[i], off_bytes(i), stream_[0]));
            }
        }
        for (int i = 0; i < n; ++i) {
            CK(cudaSetDevice(dev_[i]));
            CK(cudaStreamSynchronize(stream_[i]));
        }
        collect_stats();
        *d_tok_off = reinterpret_cast<uint64_t>(gather_off_);
        *d_tokens = reinterpret_cast<uint64_t>(gather_tokens_);
        *n_tokens = base[n];
    }

    // ---- everything else is forwarded ------------------------------------------------------------------
    void set_counting(bool on) override {
        for (auto& e : eng_) e->set_counting(on);
    }
    void set_option(const std::string& name, long long value) override {
        if (name == "output_mode" && value != 0)
            throw Error(kUnsupported, "the output stage runs on single-device tokenizers only");
        if (name == "chunk_sentences") return;  // shards are not chunked (see the constructor)
        for (auto& e : eng_) e->set_option(name, value);
    }
    void set_stream(uint64_t stream) override {
        if (stream) throw Error(kUnsupported, "a multi-device tokenizer runs on its own streams");
    }
    const float* stage_ms() const override { return stage_ms_; }
    uint64_t launch_count() const override { return launches_; }
    const uint64_t* counters() const override { return counters_; }
    void connid_counts(uint64_t* lid, uint64_t* rid, uint32_t* num_left, uint32_t* num_right) override {
        uint32_t nl = 0, nr = 0;
        eng_[0]->connid_counts(nullptr, nullptr, &nl, &nr);
        if (num_left) *num_left = nl;
        if (num_right) *num_right = nr;
        if (!lid || !rid) return;
        std::vector<uint64_t> l(nl), r(nr);
        std::fill(lid, lid + nl, 0);
        std::fill(rid, rid + nr, 0);
        for (auto& e : eng_) {
            e->connid_counts(l.data(), r.data(), nullptr, nullptr);
            for (uint32_t i = 0; i < nl; ++i) lid[i] += l[i];
            for (uint32_t i = 0; i < nr; ++i) rid[i] += r[i];
        }
    }
    uint64_t run_shard(const char*, const uint64_t*, uint64_t, int) override { throw Error(kInternal, "not a shard engine"); }
    void fetch_shard(uint64_t*, void*, uint64_t, bool) override { throw Error(kInternal, "not a shard engine"); }
    void rebase_shard(uint64_t) override { throw Error(kInternal, "not a shard engine"); }
    void shard_outputs(uint64_t*, uint64_t*) const override { throw Error(kInternal, "not a shard engine"); }
    int device() const override { return dev_[0]; }
    uint32_t token_bytes() const override { return eng_[0]->token_bytes(); }
    std::string describe() const override {
        std::string s = "{\"devices\": [";
        for (size_t i = 0; i < dev_.size(); ++i) s += (i ? ", " : "") + std::to_string(dev_[i]);
        s += "], \"dictionary_transport\": \"" + transport_ + "\", \"token_gather\": \"" +
             std::string(dev_.size() > 1 ? (nccl_ ? "ncclSend/ncclRecv" : "cudaMemcpyPeer") : "none") + "\", \"numa_pinned\": [";
        for (size_t i = 0; i < dev_.size(); ++i) s += (i ? ", " : "") + std::string(cpus_[i].empty() ? "false" : "true");
        return s + "]}";
    }

   private:
    // Contiguous shards of about equal bytes: cut[i] = first sentence of shard i, cut[n] = n_sent.
    static std::vector<uint64_t> split_by_bytes(const uint64_t* off, uint64_t n_sent, int n) {
        std::vector<uint64_t> cut(size_t(n) + 1, n_sent);
        cut[0] = 0;
        const uint64_t first = n_sent ? off[0] : 0, total = n_sent ? off[n_sent] - first : 0;
        for (int i = 1; i < n; ++i) {
            const uint64_t target = first + total / uint64_t(n) * uint64_t(i);
            const uint64_t* p = std::lower_bound(off, off + n_sent + 1, target);
            cut[i] = std::max<uint64_t>(cut[i - 1], std::min<uint64_t>(uint64_t(p - off), n_sent));
        }
        return cut;
    }

    // One host thread per device (the calling thread takes the first): each call blocks on its own device.
    template <typename F>
    void run_on_all(F&& f) {
        const int n = int(dev_.size());
        std::vector<std::string> err(n);
        std::vector<Status> st(n, kOk);
        auto body = [&](int i) {
            try {
                f(i);
            } catch (const Error& e) {
                st[i] = e.code;
                err[i] = e.what();
            } catch (const std::exception& e) {
                st[i] = kInternal;
                err[i] = e.what();
            }
        };
        std::vector<std::thread> th;
        for (int i = 1; i < n; ++i)
            th.emplace_back([&, i] {
                pin_this_thread(cpus_[i]);
                body(i);
            });
        body(0);
        for (auto& t : th) t.join();
        for (int i = 0; i < n; ++i)
            if (st[i] != kOk) throw Error(st[i], err[i]);
    }

    void collect_stats() {
        std::memset(stage_ms_, 0, sizeof stage_ms_);
        std::memset(counters_, 0, sizeof counters_);
        launches_ = 0;
        for (auto& e : eng_) {
            const float* s = e->stage_ms();
            for (int i = 0; i < kNumStages; ++i) stage_ms_[i] = std::max(stage_ms_[i], s[i]);  // devices run side by side
            const uint64_t* c = e->counters();
            for (int i = 0; i < 10; ++i) counters_[i] += c[i];
            launches_ += e->launch_count();
        }
    }

    HostResult* acquire(uint64_t n_sent, uint64_t n_tokens) {
        HostResult* r = nullptr;
        if (!free_.empty()) {
            r = free_.back();
            free_.pop_back();
        } else {
            r = new HostResult();
            pool_.push_back(r);
        }
        if ((n_sent + 1) * 8 > r->cap_off) {
            pinned_free(r->tok_off);
            r->cap_off = size_t(double((n_sent + 1) * 8) * 1.25) + 64;
            r->tok_off = static_cast<uint64_t*>(pinned_alloc(r->cap_off));
        }
        if (n_tokens * 24 > r->cap_tok) {
            pinned_free(r->tokens);
            r->cap_tok = size_t(double(n_tokens * 24) * 1.25) + 64;
            r->tokens = pinned_alloc(r->cap_tok);
        }
        r->n_sent = n_sent;
        r->n_tokens = n_tokens;
        r->has_text = false;
        if (n_sent == 0) r->tok_off[0] = 0;
        return r;
    }

    std::vector<int> dev_;
    Nccl* nccl_ = nullptr;
    std::vector<ncclComm_t> comm_;
    std::vector<cudaStream_t> stream_;
    std::vector<void*> image_;
    std::vector<std::vector<int>> cpus_;
    std::vector<std::unique_ptr<Engine>> eng_;
    std::string transport_;
    std::vector<HostResult*> pool_, free_;
    uint64_t* h_off_ = nullptr;  // pinned
    size_t h_off_cap_ = 0;
    void* gather_tokens_ = nullptr;
    void* gather_off_ = nullptr;
    size_t gather_tok_cap_ = 0, gather_off_cap_ = 0;
    float stage_ms_[kNumStages];
    uint64_t counters_[10];
    uint64_t launches_ = 0;
};

}  // namespace

std::unique_ptr<Engine> Engine::create_multi(const std::vector<int>& devices, const uint8_t* host_blob, uint64_t n_bytes,
                                             bool ignore_space, uint64_t max_grouping_len) {
    if (devices.empty()) throw Error(kInvalidArgument, "no device given");
    return std::unique_ptr<Engine>(new MultiEngine(devices, host_blob, n_bytes, ignore_space, max_grouping_len));
}

void pin_thread_to_device_numa_node(int device) { pin_this_thread(numa_cpus_of_device(device)); }

}  // namespace vbt
