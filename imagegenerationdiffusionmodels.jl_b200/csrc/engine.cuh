// Engine: owns all device state of one GPU (parameter arena, packed weights, activation
// buffers, streams, CUDA graphs, NCCL communicator) and sequences the kernels of
//   - the U-Net forward in train / inference mode  (/root/reference/src/train_brain.jl:159-179)
//   - the backward pass + Adam                      (train_brain.jl:267-272)
//   - the reverse-diffusion loop                    (/root/reference/src/generate_images.jl:174-245)
#pragma once
#include <vector>
#include <map>
#include <string>
#include <cmath>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: ranges cost ~nothing unless a profiler is attached
#include "common.cuh"
#include "kernels.cuh"
#include "igemm_simt.cuh"
#include "conv_tc.cuh"
#include "conv1_tc.cuh"
#include "l1_bwd.cuh"

namespace ddpm {

// NVTX range around the launches of one layer / phase (visible in Nsight Systems / ncu --nvtx as fwd.L3, bwd.L7, ...)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    NvtxRange(const char* prefix, int l) {
        char buf[32];
        snprintf(buf, sizeof buf, "%s%d", prefix, l);
        nvtxRangePushA(buf);
    }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

constexpr int NUM_ARRAYS = 64;
constexpr int NUM_CONV = 10;  // 3x3 convs followed by BatchNorm(relu)

// ------------------------------------------------------------------------------------ model description
struct ConvSpec {
    int cin, cout, hw;   // hw = spatial size of the output (32 or 16)
    int w, b;            // array indices of weight / bias
    int bn;              // array index of BN beta (gamma = +1, mu = +2, var = +3)
};
// constructor order of SimpleUNet (train_brain.jl:109-145); index 0 unused so that L[1]..L[10]
static const ConvSpec kConv[NUM_CONV + 1] = {
    {0, 0, 0, 0, 0, 0},
    {129, 64, 32, 0, 1, 2},     // down1.conv1
    {64, 64, 32, 6, 7, 8},      // down1.conv2 -> h1
    {64, 128, 16, 12, 13, 14},  // down2.conv1 (after MaxPool)
    {128, 128, 16, 18, 19, 20}, // down2.conv2
    {128, 128, 16, 24, 25, 26}, // mid.conv1
    {128, 128, 16, 30, 31, 32}, // mid.conv2
    {64, 64, 32, 38, 39, 40},   // up2.conv1 (after ConvTranspose, arrays 36,37)
    {64, 64, 32, 44, 45, 46},   // up2.conv2
    {128, 64, 32, 50, 51, 52},  // up1.conv1 on cat(up, h1)
    {64, 64, 32, 56, 57, 58},   // up1.conv2
};
constexpr int kUpW = 36, kUpB = 37, kFinalW = 62, kFinalB = 63;

inline void array_lengths(long long* lens) {
    int k = 0;
    auto conv = [&](int ci, int co, int kk) { lens[k++] = (long long)kk * kk * ci * co; lens[k++] = co; };
    auto bn = [&](int c) { for (int i = 0; i < 4; ++i) lens[k++] = c; };
    conv(129, 64, 3); bn(64); conv(64, 64, 3); bn(64);
    conv(64, 128, 3); bn(128); conv(128, 128, 3); bn(128);
    conv(128, 128, 3); bn(128); conv(128, 128, 3); bn(128);
    conv(128, 64, 2); conv(64, 64, 3); bn(64); conv(64, 64, 3); bn(64);
    conv(128, 64, 3); bn(64); conv(64, 64, 3); bn(64);
    conv(64, 1, 1);
}

// ------------------------------------------------------------------------------------ NCCL (dlopen'd: no link-time dependency)
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        if (lib) return;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        DDPM_CHECK(lib != nullptr, "cannot dlopen libnccl.so.2");
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        DDPM_CHECK(GetUniqueId && CommInitRank && AllReduce && CommDestroy, "libnccl is missing required symbols");
    }
    void check(ncclResult_t r, const char* what) {
        if (r != ncclSuccess) {
            std::string m = std::string(what) + " failed: " + (GetErrorString ? GetErrorString(r) : "nccl error");
            throw Error(m);
        }
    }
};
inline NcclApi& nccl() {
    static NcclApi api;
    return api;
}

// ------------------------------------------------------------------------------------ padded tensor
struct Tensor {
    void* base = nullptr;
    int C = 0;
    Geo g{};
    size_t esz = 0, bytes = 0;
    template <typename T> T* pos0() const { return reinterpret_cast<T*>(base) + (size_t)g.guard * C; }
    template <typename T> View<T> view(int c_off = 0) const { return View<T>{pos0<T>() + c_off, C}; }
    template <typename T> View<const T> cview(int c_off = 0) const { return View<const T>{pos0<T>() + c_off, C}; }
};

struct DevBuf {  // grow-only raw device buffer
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) DDPM_CUDA(cudaFree(p));
        p = nullptr; cap = 0;
        DDPM_CUDA(cudaMalloc(&p, bytes));
        cap = bytes;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct GraphEntry {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
};
// A captured training iteration: the launch sequence cut into segments at every point where a gradient bucket goes to
// NCCL.  The collectives themselves are issued eagerly between the segment launches (an NCCL call captured into a graph
// next to uncaptured calls on the same communicator dead-locked on B200 / NCCL 2.28 -- measured in round 2), everything
// else -- ~120 kernels and memsets, the peer-memory SyncBN exchanges included -- replays from the graphs.
struct TrainGraph {
    std::vector<GraphEntry> segs;
    std::vector<std::pair<int, int>> buckets;     // arena array range [a0, a1) all-reduced after segment i
    long long launches = 0;
};

// Activation set for one batch size.  Training keeps every y (pre-BN) and a (post-ReLU);
// inference aliases a small rotating set and never materialises y.
struct ActSet {
    int N = 0;
    bool training = false;
    Tensor y[NUM_CONV + 1], a[NUM_CONV + 1], p1, u;
    // gradient scratch (training only)
    Tensor g32a, g32b, gcat, g16a, g16b, gp1, gdu4;
    std::vector<void*> owned;
    DevBuf x, eps_hat;  // boundary-layout Float32 [N][H*W]
    DevBuf z;           // host-supplied sampler noise [steps][N][H*W]
    DevBuf zstep;       // device-generated noise of the current step [N][H*W]
    DevBuf rng;         // [seed, first_index] of the chunk this set is currently sampling (read by its graphs)
    // captured reverse loops, keyed by (t_start, zmode); they bake in this set's pointers
    std::map<std::pair<int, int>, GraphEntry> graphs;
    // training only: device staging of one step's inputs (host batches or dataset indices), q_sample output, loss
    // gradient and the first layer's backward scratch -- owned by the set so that a captured step never sees a
    // re-allocated pointer
    DevBuf x0, eps, ts, idx, xt, deps, Ccls, S;
    // captured training iterations, keyed by gather*2 + update (the first call of a key runs eagerly and warms every
    // lazily initialised resource, the second one is captured, later ones replay)
    std::map<int, TrainGraph> train_graphs;
    std::map<int, int> train_calls;
    long long last_use = 0;
    static void destroy(GraphEntry& g) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
        g = GraphEntry();
    }
    void drop_graphs() {
        for (auto& kv : graphs) destroy(kv.second);
        graphs.clear();
        for (auto& kv : train_graphs)
            for (auto& g : kv.second.segs) destroy(g);
        train_graphs.clear();
        train_calls.clear();
    }
};

enum class Mode { Infer, Train };

struct Engine {
    int T, D, H, W, prec, dev;
    int HW;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_bucket[5] = {}, ev_comm_done = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    // Sampling can run consecutive chunks on alternating streams (chunks are independent).  Measured on B200
    // (round 1): no gain -- 4096 images, chunk 512: 1 stream 1786 img/s, 2 streams 1754; chunk 256 x 2 streams 1704 --
    // the persistent conv kernels already own every SM, so this stays an option (default 1).
    static constexpr int MAX_SAMPLE_STREAMS = 2;
    cudaStream_t sample_streams[MAX_SAMPLE_STREAMS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_SAMPLE_STREAMS] = {};
    long long lens[NUM_ARRAYS], offs[NUM_ARRAYS + 1];
    long long n_params = 0;

    // tables
    std::vector<float> h_beta, h_acum, h_pe, h_samp;  // host copies
    float *d_sqrt_ac = nullptr, *d_sqrt_1mac = nullptr, *d_pe = nullptr;

    // parameters
    float *P = nullptr, *G = nullptr, *M1 = nullptr, *M2 = nullptr;
    float eta = 1e-4f, b1 = 0.9f, b2 = 0.999f, aeps = 1e-8f;
    // Optimiser state that changes every step lives on the device (TrainState, kernels.cuh): the bias-correction
    // powers beta^t, the non-finite-gradient flag of the current step and the skipped / applied step counters.  The
    // host never needs them to enqueue a step, so a whole training iteration can be replayed from a CUDA graph.
    TrainState* d_tstate = nullptr;
    tc::WgScratch wg;             // partial-sum scratch of the tcgen05 weight-gradient kernels (per engine)
    long long opt_loss_scale_log2 = 0;   // extra power-of-two factor on the static loss scale (tests trip the overflow guard with it)
    // Static loss scale of the 16-bit gradient tensors: d(loss)/d(eps_hat) is multiplied by
    // S = B_global*H*W/8 (so it is (eps_hat-eps)/4, O(1)) and every FP32 gradient written to the arena is
    // multiplied by 1/S.  Powers of two when B is: exact.  Keeps FP16 gradients ~3 decades below overflow
    // and the bulk above the subnormal range (measured ranges in DESIGN.md).  1 in FP32 mode.
    float grad_scale(int B) const {
        const float base = (prec == 0 || prec == 3) ? 1.f : (float)B * (float)world * (float)HW / 8.f;
        return std::ldexp(base, (int)opt_loss_scale_log2);
    }

    // packed weights
    void* Wf[NUM_CONV + 1] = {};   // [cout][9][cin]  (TA)  -- L1 unused
    void* Wd[NUM_CONV + 1] = {};   // [cin][9][cout]  (TG)  -- L1 unused
    void* Wfi[NUM_CONV + 1] = {};  // Wf with the inference BatchNorm scale folded in (TA)
    void *Wt = nullptr, *Wtd = nullptr;  // convT fwd (TA) [256][128], dgrad (TG) [128][256]
    float *Wimg = nullptr, *Wemb = nullptr;  // first conv, FP32
    float *Ptab = nullptr, *Ecls = nullptr;  // [T][9][64] embedding contributions
    bool ecls_valid = false, infer_affine_valid = false;

    // per-BN-layer device vectors
    float *inf_scale[NUM_CONV + 1] = {}, *inf_shift[NUM_CONV + 1] = {};
    float *tr_mean[NUM_CONV + 1] = {}, *tr_istd[NUM_CONV + 1] = {}, *tr_scale[NUM_CONV + 1] = {}, *tr_shift[NUM_CONV + 1] = {};
    float *bw_mg[NUM_CONV + 1] = {}, *bw_mgx[NUM_CONV + 1] = {};
    double* sums = nullptr;       // [NUM_CONV+1][3*128] forward (sum, sumsq) / backward (g, g*xhat, dy)
    double* sums_g = nullptr;     // all-reduced copies (SyncBN)
    double* misc_sums = nullptr;  // [0]=loss, [8..72]=dwf, [72]=dbf, [128..192]=dbT
    double* l1_acc = nullptr;     // [576] image-channel weight gradient of the first conv, summed over the batch

    // activation sets: a two-entry cache of training sets (an epoch of the reference alternates 64- and 52-image
    // batches, train_brain.jl:201-202) and a small cache of inference sets, both keyed by batch size
    std::map<int, ActSet*> train_sets;
    ActSet* last_train_set = nullptr;
    long long use_clock = 0;
    std::map<int, ActSet*> infer_sets;
    DevBuf d_x0, d_eps, d_xt, d_ts, d_dataset, d_sample_out, d_u8;     // q_sample API staging, dataset, sampler output
    unsigned long long* d_trng = nullptr;     // [seed, first_index, step] of the current device-drawn training step
    long long dataset_n = 0;
    unsigned long long* d_rng = nullptr;  // [seed, first_index]

    // communicator
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, sync_bn = 0;
    // peer-memory mailboxes for the tiny SyncBN all-reduces (kernels.cuh: XReduce); xr_ok once every peer is mapped
    XReduce xr{};
    double* xr_own = nullptr;
    void* xr_mapped[XR_MAX_WORLD] = {};
    bool xr_ok = false;
    std::string xr_note = "not initialised";
    long long opt_bn_p2p = 1;      // SyncBN statistics over the peer-memory mailboxes (0: NCCL all-reduce per layer)
    long long opt_dp_skip = 0;     // TIMING ONLY (results become wrong): bit 0 skips the gradient all-reduces, bit 1 the
                                   // SyncBN all-reduces -- used to split a data-parallel step into compute / exposed comm
    void init_peer_mailboxes();

    // options / counters
    long long opt_sample_streams = 1;
    long long opt_train_reverse = 1; // BatchNorm apply / backward kernels walk their tensors from the end (L2-resident tail first)
    long long opt_bnbwd_blocks = 4; // resident blocks per SM of the second BatchNorm-backward pass
    long long opt_conv1_tc = 2;    // sampler: first conv on tensor cores (hi/lo split operands) when the batch shares one timestep;
                                   // 2 = (timestep, border class) constants folded into the contraction as well, 1 = added in the epilogue
    // images per captured reverse-loop graph.  1300 images fill the persistent conv kernels' tile rounds exactly
    // (32x32 layers: 77.0 rounds of 74 CTA-pair tiles, 16x16 layers: 21.0) -- 512 left the 16x16 layers at 92 %
    long long opt_sample_chunk = 1300, opt_use_graph = 1, opt_conv_impl = 0 /*0 auto, 1 simt, 2 tc*/, opt_fuse_final = 1;
    long long opt_fuse_bn = 1;              // BatchNorm reductions inside the tcgen05 conv / data-gradient epilogues
    long long opt_train_graph = 1;          // replay the training iteration from a CUDA graph (per set and step kind)
    long long cnt_launches = 0;

    Engine(int T_, int D_, int H_, int W_, int prec_, int dev_);
    ~Engine();

    // ---- helpers
    size_t esz_a() const { return (prec == 0 || prec == 3) ? 4 : 2; }
    size_t esz_g() const { return (prec == 0 || prec == 3) ? 4 : 2; }
    float* arr(int k) const { return P + offs[k]; }
    float* garr(int k) const { return G + offs[k]; }
    double* lsum(int l) const { return sums + (size_t)l * 384; }
    double* gsum(int l) const { return (sync_bn && comm ? sums_g : sums) + (size_t)l * 384; }
    bool use_tc() const {
        if (opt_conv_impl == 1) return false;
        if (opt_conv_impl == 2) return true;
        return prec != 0 && tc::available();
    }

    void reset_train_state();
    void set_default_tables();
    void upload_tables();
    void alloc_tensor(ActSet& s, Tensor& t, int N, int hw, int C, size_t esz);
    void build_set(ActSet& s, int N, bool training);
    void free_set(ActSet& s);
    ActSet& get_set(int N, bool training, int slot = 0);

    template <typename TA, typename TG> void pack_weights_t();
    template <typename TA, typename TG> void pack_infer_weights_t();
    void pack_weights();
    void prepare_ecls();
    void prepare_infer_affine();

    template <typename TA, typename TG>
    void conv3(const Tensor& s0, const Tensor* s1, int l, Tensor& out, bool infer_weights, const float* shift, int relu,
               double* stats, int rev = 0);
    template <typename TA, typename TG>
    void dgrad3(ActSet& s, const Tensor& dy, int l, Tensor& out, int out_c_total, int bn_layer = 0, bool* bn_done = nullptr);
    tc::BnFuse bn_bwd_fuse(ActSet& s, int bn_layer);
    template <typename TA, typename TG> void forward_t(ActSet& s, const float* x_dev, const int* ts_dev, int t_fixed, Mode mode, bool update_running,
                                                       bool skip_last = false);
    void forward(ActSet& s, const float* x_dev, const int* ts_dev, int t_fixed, Mode mode, bool update_running);
    template <typename TA, typename TG> void final_conv_t(ActSet& s, float* eps_hat_dev);
    template <typename TA, typename TG> void layer10_fallback(ActSet& s) {
        conv3<TA, TG>(s.a[9], nullptr, 10, s.a[10], true, inf_shift[10], 1, nullptr);
    }
    template <typename TA, typename TG> void backward_t(ActSet& s, const float* xt_dev, const int* ts_dev, const float* deps_dev, float alpha);
    void allreduce_sums(double* local, double* global, int n);

    void train_enqueue(ActSet& s, bool gather, bool device_draws, bool update);
    void train_core(ActSet& s, bool gather, bool device_draws, bool update, float* loss_out_host);
    TrainGraph* capturing = nullptr;        // non-null while train_core records a training iteration
    void seg_begin() { DDPM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal)); }
    void seg_end() {
        GraphEntry ge;
        DDPM_CUDA(cudaStreamEndCapture(stream, &ge.graph));
        DDPM_CUDA(cudaGraphInstantiate(&ge.exec, ge.graph, 0));
        capturing->segs.push_back(ge);
    }
    void issue_bucket(int a0, int a1, int k) {
        cudaEvent_t ev = ev_bucket[k % 5];
        DDPM_CUDA(cudaEventRecord(ev, stream));
        DDPM_CUDA(cudaStreamWaitEvent(comm_stream, ev, 0));
        nccl().check(nccl().AllReduce(G + offs[a0], G + offs[a0], (size_t)(offs[a1] - offs[a0]), ncclFloat32, ncclSum, comm,
                                      comm_stream), "ncclAllReduce(gradient bucket)");
    }
    void join_comm() {
        DDPM_CUDA(cudaEventRecord(ev_comm_done, comm_stream));
        DDPM_CUDA(cudaStreamWaitEvent(stream, ev_comm_done, 0));
    }
    void drop_train_graphs() { for (auto& kv : train_sets) kv.second->drop_graphs(); }
    void drop_all_graphs() { drop_train_graphs(); for (auto& kv : infer_sets) kv.second->drop_graphs(); }
    void sample_chunk(ActSet& s, bool host_z, unsigned long long seed, long long first_index, int t_start);
    template <typename TA, typename TG> void sample_steps_t(ActSet& s, float* x_dev, const float* z_dev, int N, int t_start);
};

// ===================================================================================== implementation

#define DDPM_DISPATCH(prec, ...)                                                         \
    do {                                                                                 \
        if ((prec) == 0 || (prec) == 3) { using TA = float; using TG = float; __VA_ARGS__; } \
        else if ((prec) == 1) { using TA = __half; using TG = __half; __VA_ARGS__; }     \
        else { using TA = __nv_bfloat16; using TG = __nv_bfloat16; __VA_ARGS__; }        \
    } while (0)

inline Engine::Engine(int T_, int D_, int H_, int W_, int prec_, int dev_)
    : T(T_), D(D_), H(H_), W(W_), prec(prec_), dev(dev_), HW(H_ * W_) {
    DDPM_CHECK(T >= 2 && T <= 100000, "T out of range");
    DDPM_CHECK(D == 128, "only D=128 (the reference's embedding width) is supported");
    DDPM_CHECK(H == 32 && W == 32, "only 32x32 images (the reference's data) are supported");
    DDPM_CHECK(prec >= 0 && prec <= 3, "precision must be 0 (fp32), 1 (fp16), 2 (bf16) or 3 (tf32)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw Error("no CUDA device: libddpm has no CPU fallback");
    DDPM_CHECK(dev >= 0 && dev < ndev, "device index out of range");
    DDPM_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    DDPM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) throw Error("libddpm is built for sm_100a (B200) only");
    DDPM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    DDPM_CUDA(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
    for (auto& ev : ev_bucket) DDPM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    DDPM_CUDA(cudaEventCreateWithFlags(&ev_comm_done, cudaEventDisableTiming));
    DDPM_CUDA(cudaEventCreate(&ev_t0));
    DDPM_CUDA(cudaEventCreate(&ev_t1));
    sample_streams[0] = stream;
    for (int i = 1; i < MAX_SAMPLE_STREAMS; ++i) DDPM_CUDA(cudaStreamCreateWithFlags(&sample_streams[i], cudaStreamNonBlocking));
    DDPM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    for (auto& ev : ev_join) DDPM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

    array_lengths(lens);
    offs[0] = 0;
    for (int k = 0; k < NUM_ARRAYS; ++k) {
        // keep every array 16-byte aligned inside the arena
        offs[k + 1] = offs[k] + ((lens[k] + 3) / 4) * 4;
    }
    n_params = offs[NUM_ARRAYS];
    // every initialisation below is enqueued on `stream` (a non-blocking stream is NOT ordered after legacy-stream
    // cudaMemset / cudaMemcpy); the constructor synchronises once at the end
    auto zalloc = [&](float** p, size_t n) {
        DDPM_CUDA(cudaMalloc(p, n * sizeof(float)));
        DDPM_CUDA(cudaMemsetAsync(*p, 0, n * sizeof(float), stream));
    };
    zalloc(&P, n_params); zalloc(&G, n_params); zalloc(&M1, n_params); zalloc(&M2, n_params);
    zalloc(&d_sqrt_ac, T); zalloc(&d_sqrt_1mac, T); zalloc(&d_pe, (size_t)T * D);
    zalloc(&Wimg, 9 * 64); zalloc(&Wemb, (size_t)9 * 64 * D);
    zalloc(&Ptab, (size_t)T * 576); zalloc(&Ecls, (size_t)T * 576);
    for (int l = 1; l <= NUM_CONV; ++l) {
        int C = kConv[l].cout;
        zalloc(&inf_scale[l], C); zalloc(&inf_shift[l], C);
        zalloc(&tr_mean[l], C); zalloc(&tr_istd[l], C); zalloc(&tr_scale[l], C); zalloc(&tr_shift[l], C);
        zalloc(&bw_mg[l], C); zalloc(&bw_mgx[l], C);
        if (l >= 2) {
            size_t n = (size_t)9 * kConv[l].cin * kConv[l].cout;
            DDPM_CUDA(cudaMalloc(&Wf[l], n * esz_a()));
            DDPM_CUDA(cudaMalloc(&Wd[l], n * esz_g()));
            DDPM_CUDA(cudaMalloc(&Wfi[l], n * esz_a()));
        }
    }
    DDPM_CUDA(cudaMalloc(&Wt, (size_t)4 * 128 * 64 * esz_a()));
    DDPM_CUDA(cudaMalloc(&Wtd, (size_t)4 * 128 * 64 * esz_g()));
    DDPM_CUDA(cudaMalloc(&sums, sizeof(double) * 384 * (NUM_CONV + 1)));
    DDPM_CUDA(cudaMalloc(&sums_g, sizeof(double) * 384 * (NUM_CONV + 1)));
    DDPM_CUDA(cudaMalloc(&misc_sums, sizeof(double) * 256));
    DDPM_CUDA(cudaMalloc(&l1_acc, sizeof(double) * 576));
    DDPM_CUDA(cudaMalloc(&d_rng, 2 * sizeof(unsigned long long)));
    DDPM_CUDA(cudaMalloc(&d_trng, 4 * sizeof(unsigned long long)));
    DDPM_CUDA(cudaMalloc(&d_tstate, sizeof(TrainState)));
    reset_train_state();
    tc::init();
    set_default_tables();
    // Flux default initialisation of the BatchNorm state so an un-loaded handle is well defined:
    // gamma = 1, var = 1; conv weights stay zero until ddpm_set_weights.
    const std::vector<float> ones(128, 1.f);
    for (int l = 1; l <= NUM_CONV; ++l) {
        int C = kConv[l].cout;
        DDPM_CUDA(cudaMemcpyAsync(arr(kConv[l].bn + 1), ones.data(), C * 4, cudaMemcpyHostToDevice, stream));
        DDPM_CUDA(cudaMemcpyAsync(arr(kConv[l].bn + 3), ones.data(), C * 4, cudaMemcpyHostToDevice, stream));
    }
    pack_weights();
    DDPM_CUDA(cudaStreamSynchronize(stream));
}

inline void Engine::reset_train_state() {
    TrainState ts{};
    ts.bt1 = b1; ts.bt2 = b2;
    DDPM_CUDA(cudaMemcpyAsync(d_tstate, &ts, sizeof ts, cudaMemcpyHostToDevice, stream));
    DDPM_CUDA(cudaStreamSynchronize(stream));
}

inline Engine::~Engine() {
    cudaSetDevice(dev);
    cudaDeviceSynchronize();
    for (int r = 0; r < XR_MAX_WORLD; ++r)
        if (xr_mapped[r]) cudaIpcCloseMemHandle(xr_mapped[r]);
    if (xr_own) cudaFree(xr_own);
    if (xr.epoch) cudaFree(xr.epoch);
    if (comm) nccl().CommDestroy(comm);
    for (auto& kv : train_sets) { free_set(*kv.second); delete kv.second; }
    train_sets.clear();
    for (auto& kv : infer_sets) { free_set(*kv.second); delete kv.second; }
    infer_sets.clear();
    float* fl[] = {P, G, M1, M2, d_sqrt_ac, d_sqrt_1mac, d_pe, Wimg, Wemb, Ptab, Ecls};
    for (float* p : fl) cudaFree(p);
    for (int l = 1; l <= NUM_CONV; ++l) {
        float* v[] = {inf_scale[l], inf_shift[l], tr_mean[l], tr_istd[l], tr_scale[l], tr_shift[l], bw_mg[l], bw_mgx[l]};
        for (float* p : v) cudaFree(p);
        cudaFree(Wf[l]); cudaFree(Wd[l]); cudaFree(Wfi[l]);
    }
    cudaFree(Wt); cudaFree(Wtd); cudaFree(sums); cudaFree(sums_g); cudaFree(misc_sums); cudaFree(l1_acc); cudaFree(d_rng); cudaFree(d_tstate);
    wg.release();
    cudaFree(d_trng);
    DevBuf* bufs[] = {&d_x0, &d_eps, &d_xt, &d_ts, &d_dataset, &d_sample_out, &d_u8};
    for (DevBuf* b : bufs) b->release();
    for (auto& ev : ev_bucket) cudaEventDestroy(ev);
    cudaEventDestroy(ev_comm_done);
    cudaEventDestroy(ev_t0); cudaEventDestroy(ev_t1);
    for (int i = 1; i < MAX_SAMPLE_STREAMS; ++i) cudaStreamDestroy(sample_streams[i]);
    cudaEventDestroy(ev_fork);
    for (auto& ev : ev_join) cudaEventDestroy(ev);
    cudaStreamDestroy(stream);
    cudaStreamDestroy(comm_stream);
}

// The library's own restatement of the script constants (train_brain.jl:20-24,54-62); the host
// normally overrides it with ddpm_set_tables so the tables are bit-exact with the host language.
inline void Engine::set_default_tables() {
    h_beta.resize(T); h_acum.resize(T); h_pe.resize((size_t)T * D);
    const double b0 = (double)1e-4f, b1d = (double)0.02f;
    float prod = 1.f;
    for (int i = 0; i < T; ++i) {
        h_beta[i] = (float)(b0 + i * (b1d - b0) / (T - 1));
        float alpha = 1.f - h_beta[i];
        prod = (i == 0) ? alpha : prod * alpha;
        h_acum[i] = prod;
    }
    const double neg_log = -(double)logf(1e4f);
    for (int t = 1; t <= T; ++t)
        for (int i = 1; i <= D / 2; ++i) {
            double div = exp(neg_log * (2.0 * (i - 1) / (D - 1)));
            h_pe[(size_t)(t - 1) * D + 2 * i - 2] = (float)sin(t * div);
            h_pe[(size_t)(t - 1) * D + 2 * i - 1] = (float)cos(t * div);
        }
    upload_tables();
}

inline void Engine::upload_tables() {
    // per-step scalars exactly in the order generate_images.jl:186-208 writes them (Float32)
    std::vector<float> sa(T), sb(T);
    h_samp.assign((size_t)T * 4, 0.f);
    for (int t = 1; t <= T; ++t) {
        volatile float a_t = h_acum[t - 1];
        volatile float a_prev = t > 1 ? h_acum[t - 2] : 1.f;
        volatile float beta_t = 1.f - a_t;
        volatile float beta_prev = 1.f - a_prev;
        volatile float one_m = 1.f - a_t;
        volatile float num = beta_prev * one_m;
        volatile float pv = num / one_m;
        h_samp[(t - 1) * 4 + 0] = sqrtf(beta_t);
        h_samp[(t - 1) * 4 + 1] = sqrtf(a_t);
        h_samp[(t - 1) * 4 + 2] = sqrtf(a_prev);
        h_samp[(t - 1) * 4 + 3] = sqrtf(pv);
        sa[t - 1] = sqrtf(a_t);
        volatile float om = 1.f - a_t;
        sb[t - 1] = sqrtf(om);
    }
    DDPM_CUDA(cudaMemcpyAsync(d_sqrt_ac, sa.data(), T * 4, cudaMemcpyHostToDevice, stream));
    DDPM_CUDA(cudaMemcpyAsync(d_sqrt_1mac, sb.data(), T * 4, cudaMemcpyHostToDevice, stream));
    DDPM_CUDA(cudaMemcpyAsync(d_pe, h_pe.data(), (size_t)T * D * 4, cudaMemcpyHostToDevice, stream));
    DDPM_CUDA(cudaStreamSynchronize(stream));      // sa / sb are locals
    ecls_valid = false;
    for (auto& kv : infer_sets) kv.second->drop_graphs();  // scalars are baked into captured graphs
}

// ------------------------------------------------------------------------------------ buffers
inline void Engine::alloc_tensor(ActSet& s, Tensor& t, int N, int hw, int C, size_t esz) {
    t.C = C; t.g = Geo::make(N, hw, hw); t.esz = esz;
    t.bytes = (size_t)t.g.alloc_positions() * C * esz;
    DDPM_CUDA(cudaMalloc(&t.base, t.bytes));
    DDPM_CUDA(cudaMemsetAsync(t.base, 0, t.bytes, stream));
    s.owned.push_back(t.base);
}

inline void Engine::free_set(ActSet& s) {
    s.drop_graphs();
    for (void* p : s.owned) cudaFree(p);
    s.owned.clear();
    DevBuf* bufs[] = {&s.x, &s.eps_hat, &s.z, &s.zstep, &s.rng, &s.x0, &s.eps, &s.ts, &s.idx, &s.xt, &s.deps, &s.Ccls, &s.S};
    for (DevBuf* b : bufs) b->release();
    s = ActSet();
}

inline void Engine::build_set(ActSet& s, int N, bool training) {
    free_set(s);
    s.N = N; s.training = training;
    const size_t ea = esz_a(), eg = esz_g();
    if (training) {
        for (int l = 1; l <= NUM_CONV; ++l) {
            alloc_tensor(s, s.y[l], N, kConv[l].hw, kConv[l].cout, ea);
            alloc_tensor(s, s.a[l], N, kConv[l].hw, kConv[l].cout, ea);
        }
        alloc_tensor(s, s.p1, N, 16, 64, ea);
        alloc_tensor(s, s.u, N, 32, 64, ea);
        alloc_tensor(s, s.g32a, N, 32, 64, eg);
        alloc_tensor(s, s.g32b, N, 32, 64, eg);
        alloc_tensor(s, s.gcat, N, 32, 128, eg);
        alloc_tensor(s, s.g16a, N, 16, 128, eg);
        alloc_tensor(s, s.g16b, N, 16, 128, eg);
        alloc_tensor(s, s.gp1, N, 16, 64, eg);
        alloc_tensor(s, s.gdu4, N, 16, 256, eg);
    } else {
        Tensor i32[4], i16[2];
        for (auto& t : i32) alloc_tensor(s, t, N, 32, 64, ea);
        for (auto& t : i16) alloc_tensor(s, t, N, 16, 128, ea);
        alloc_tensor(s, s.p1, N, 16, 64, ea);
        s.a[1] = i32[0]; s.a[2] = i32[1];
        s.a[3] = i16[0]; s.a[4] = i16[1]; s.a[5] = i16[0]; s.a[6] = i16[1];
        s.u = i32[0]; s.a[7] = i32[2]; s.a[8] = i32[0]; s.a[9] = i32[2]; s.a[10] = i32[3];
    }
    s.x.ensure((size_t)N * HW * 4);
    s.eps_hat.ensure((size_t)N * HW * 4);
    if (!training) { s.zstep.ensure((size_t)N * HW * 4); s.rng.ensure(2 * sizeof(unsigned long long)); }
    if (training) {
        const size_t img = (size_t)N * HW * 4;
        s.x0.ensure(img); s.eps.ensure(img); s.xt.ensure(img); s.deps.ensure(img);
        s.ts.ensure((size_t)N * 4); s.idx.ensure((size_t)N * 4);
        s.Ccls.ensure((size_t)N * 576 * 4); s.S.ensure((size_t)N * 576 * 4);
        if (use_tc()) tc::wg_reserve(wg);          // no allocation may happen inside a captured step
    }
}

inline ActSet& Engine::get_set(int N, bool training, int slot) {
    if (training) {
        auto it = train_sets.find(N);
        if (it == train_sets.end()) {
            DDPM_CUDA(cudaStreamSynchronize(stream));
            if (train_sets.size() >= 2) {             // evict the least recently used set
                auto victim = train_sets.begin();
                for (auto jt = train_sets.begin(); jt != train_sets.end(); ++jt)
                    if (jt->second->last_use < victim->second->last_use) victim = jt;
                if (last_train_set == victim->second) last_train_set = nullptr;
                free_set(*victim->second);
                delete victim->second;
                train_sets.erase(victim);
            }
            ActSet* ns = new ActSet();
            build_set(*ns, N, true);
            DDPM_CUDA(cudaStreamSynchronize(stream));
            it = train_sets.emplace(N, ns).first;
        }
        it->second->last_use = ++use_clock;
        last_train_set = it->second;
        return *it->second;
    }
    const int key = N * 8 + slot;   // one set per (batch size, concurrent-stream slot)
    auto it = infer_sets.find(key);
    if (it != infer_sets.end()) return *it->second;
    DDPM_CUDA(cudaDeviceSynchronize());
    if (infer_sets.size() >= 8) {  // bounded cache: drop the smallest-batch set
        auto victim = infer_sets.begin();
        free_set(*victim->second);
        delete victim->second;
        infer_sets.erase(victim);
    }
    ActSet* s = new ActSet();
    build_set(*s, N, false);
    DDPM_CUDA(cudaDeviceSynchronize());   // the zero-fill ran on whichever stream is current; chunk streams differ
    infer_sets[key] = s;
    return *s;
}

// ------------------------------------------------------------------------------------ derived weights
template <typename TA, typename TG>
void Engine::pack_weights_t() {
    PackJobs J{};
    long long nmax = 0;
    for (int l = 2; l <= NUM_CONV; ++l) {
        const ConvSpec& c = kConv[l];
        J.w[l - 2] = arr(c.w); J.cin[l - 2] = c.cin; J.cout[l - 2] = c.cout;
        J.out_f[l - 2] = Wf[l]; J.out_d[l - 2] = Wd[l]; J.row_scale[l - 2] = nullptr;
        J.tf32_round = (prec == 3) ? 1 : 0;
        nmax = std::max(nmax, 9LL * c.cin * c.cout);
    }
    pack_conv3_batch_kernel<TA, TG><<<dim3(cdiv(nmax, 256), NUM_CONV - 1, 2), 256, 0, stream>>>(J);
    cnt_launches += 1;
    long long nt = 4LL * 128 * 64;
    pack_up2_kernel<TA><<<cdiv(nt, 256), 256, 0, stream>>>(arr(kUpW), 128, 64, 0, (TA*)Wt);
    pack_up2_kernel<TG><<<cdiv(nt, 256), 256, 0, stream>>>(arr(kUpW), 128, 64, 1, (TG*)Wtd);
    long long n1 = 9LL * 64 * (D + 1);
    pack_l1_kernel<<<cdiv(n1, 256), 256, 0, stream>>>(arr(0), D, 64, Wimg, Wemb);
    DDPM_LAUNCH_CHECK();
    cnt_launches += 3;
}

inline void Engine::pack_weights() {
    DDPM_DISPATCH(prec, (pack_weights_t<TA, TG>()));
    ecls_valid = false;
    infer_affine_valid = false;
}

// Ptab[t][tap][co] = sum_c pe[t][c] * Wemb[tap][co][c]  (500 x 128 x 576 GEMM), then border-class sums
inline void Engine::prepare_ecls() {
    if (ecls_valid) return;
    View<const float> pe{d_pe, D}, none{nullptr, 0};
    EpiPlain<float> epi{Ptab, 576, (long long)T, nullptr};
    launch_igemm_simt<float, float>(stream, pe, D, none, 0, Wemb, 576, 1, (long long)T, MapId{(long long)T}, epi);
    emb_class_sums_kernel<<<cdiv((long long)T * 576, 256), 256, 0, stream>>>(Ptab, Ecls, T, 64);
    DDPM_LAUNCH_CHECK();
    cnt_launches += 2;
    ecls_valid = true;
}

template <typename TA, typename TG>
void Engine::pack_infer_weights_t() {
    PackJobs J{};
    long long nmax = 0;
    for (int l = 2; l <= NUM_CONV; ++l) {
        const ConvSpec& c = kConv[l];
        J.w[l - 2] = arr(c.w); J.cin[l - 2] = c.cin; J.cout[l - 2] = c.cout;
        J.out_f[l - 2] = Wfi[l]; J.out_d[l - 2] = nullptr; J.row_scale[l - 2] = inf_scale[l];
        J.tf32_round = (prec == 3) ? 1 : 0;
        nmax = std::max(nmax, 9LL * c.cin * c.cout);
    }
    pack_conv3_batch_kernel<TA, TG><<<dim3(cdiv(nmax, 256), NUM_CONV - 1, 1), 256, 0, stream>>>(J);
    cnt_launches += 1;
    DDPM_LAUNCH_CHECK();
}

// Inference: BatchNorm with running statistics is a per-channel affine map; its scale is folded into
// the packed weights (Wfi) and its shift into the conv epilogue.
inline void Engine::prepare_infer_affine() {
    if (infer_affine_valid) return;
    for (int l = 1; l <= NUM_CONV; ++l) {
        const ConvSpec& c = kConv[l];
        bn_inference_affine_kernel<<<1, 128, 0, stream>>>(arr(c.bn + 1), arr(c.bn), arr(c.bn + 2), arr(c.bn + 3), arr(c.b),
                                                          inf_scale[l], inf_shift[l], c.cout, 1e-5f);
    }
    DDPM_LAUNCH_CHECK();
    cnt_launches += NUM_CONV;
    DDPM_DISPATCH(prec, (pack_infer_weights_t<TA, TG>()));
    infer_affine_valid = true;
}

// ------------------------------------------------------------------------------------ convolutions
// Conv((3,3), cin=>cout, pad=1) of layer l on s0 (and s1 concatenated along channels)
template <typename TA, typename TG>
void Engine::conv3(const Tensor& s0, const Tensor* s1, int l, Tensor& out, bool infer_weights, const float* shift,
                   int relu, double* stats, int rev) {
    const void* weights = infer_weights ? Wfi[l] : Wf[l];
    const ConvSpec& c = kConv[l];
    const Geo& g = out.g;
    int C0 = s0.C, C1 = s1 ? s1->C : 0;
    DDPM_CHECK(C0 + C1 == c.cin && out.C == c.cout, "conv3: channel mismatch");
    if (use_tc()) {
        // train-mode BatchNorm statistics of the stored (rounded) y: fused into the conv epilogue when the launched
        // variant supports it, otherwise one streaming reduction over y
        tc::BnFuse bf{};
        bf.mode = (stats && opt_fuse_bn) ? 1 : 0; bf.sums = stats; bf.C = c.cout;
        bool fused = false;
        bool ok = tc::conv3x3<TA, TA>(stream, s0.pos0<TA>(), C0, s1 ? s1->pos0<TA>() : nullptr, C1, (const TA*)weights, c.cout,
                                      out.pos0<TA>(), g, shift, relu, &bf, &fused, rev);
        if (ok) {
            cnt_launches += 1;
            if (stats && !fused) {
                bn_reduce_linear_kernel<TA, TA, 0><<<lin_reduce_blocks(g.npos, tc::state().num_sms), 256, 0, stream>>>(
                    out.cview<TA>(), out.cview<TA>(), g.npos, c.cout, nullptr, nullptr, nullptr, nullptr, stats);
                DDPM_LAUNCH_CHECK();
                cnt_launches += 1;
            }
            return;
        }
    }
    MapConv3 map{g.Wp, -(long long)g.guard, g.npos + g.guard};
    EpiConv<TA> epi{out.view<TA>(), g, nullptr, shift, relu, stats};
    View<const TA> v1 = s1 ? s1->cview<TA>() : View<const TA>{nullptr, 0};
    launch_igemm_simt<TA, TA>(stream, s0.cview<TA>(), C0, v1, C1, (const TA*)weights, c.cout, 9, g.npos, map, epi);
    cnt_launches += 1;
}

// parameters of the first BatchNorm-backward pass of layer `bn_layer` for the kernel that PRODUCES its input gradient
// da = dL/d(a_l):  sum g, sum g*xhat over the tile it just wrote (see tc::TcParams::stats)
inline tc::BnFuse Engine::bn_bwd_fuse(ActSet& s, int bn_layer) {
    tc::BnFuse bf{};
    const ConvSpec& b = kConv[bn_layer];
    bf.mode = 2; bf.sums = lsum(bn_layer); bf.C = b.cout; bf.nch = b.cout;
    bf.y = s.y[bn_layer].base ? (const char*)s.y[bn_layer].base + (size_t)s.y[bn_layer].g.guard * b.cout * s.y[bn_layer].esz : nullptr;
    bf.y_cs = b.cout;
    bf.scale = tr_scale[bn_layer]; bf.shift = tr_shift[bn_layer]; bf.mean = tr_mean[bn_layer]; bf.istd = tr_istd[bn_layer];
    return bf;
}

// data gradient of layer l: out[p][ci] = sum_tap sum_co dy[p - shift(tap)][co] * W[co][tap][ci]
// bn_layer != 0: `out` (its first cout(bn_layer) channels) is the gradient w.r.t. the BatchNorm-ReLU output of that
// layer; the epilogue then also accumulates the first pass of that BatchNorm's backward (*bn_done reports it).
template <typename TA, typename TG>
void Engine::dgrad3(ActSet& s, const Tensor& dy, int l, Tensor& out, int out_c_total, int bn_layer, bool* bn_done) {
    NvtxRange r("bwd.dgrad.L", l);
    const ConvSpec& c = kConv[l];
    const Geo& g = out.g;
    if (bn_done) *bn_done = false;
    DDPM_CHECK(dy.C == c.cout && out.C == out_c_total && out_c_total == c.cin, "dgrad3: channel mismatch");
    if (use_tc()) {
        tc::BnFuse bf{};
        if (bn_layer && opt_fuse_bn) bf = bn_bwd_fuse(s, bn_layer);
        bool ok = tc::conv3x3<TG, TG>(stream, dy.pos0<TG>(), c.cout, nullptr, 0, (const TG*)Wd[l], c.cin, out.pos0<TG>(), g,
                                     nullptr, 0, &bf, bn_done);
        if (ok) {
            cnt_launches += 1;
            return;
        }
    }
    MapConv3 map{g.Wp, -(long long)g.guard, g.npos + g.guard};
    EpiConv<TG> epi{out.view<TG>(), g, nullptr, nullptr, 0, nullptr};
    launch_igemm_simt<TG, TG>(stream, dy.cview<TG>(), c.cout, View<const TG>{nullptr, 0}, 0, (const TG*)Wd[l], c.cin, 9,
                              g.npos, map, epi);
    cnt_launches += 1;
}

// ------------------------------------------------------------------------------------ forward
template <typename TA, typename TG>
void Engine::forward_t(ActSet& s, const float* x_dev, const int* ts_dev, int t_fixed, Mode mode, bool update_running,
                       bool skip_last) {
    const int N = s.N;
    const bool train = mode == Mode::Train;
    prepare_ecls();
    if (!train) prepare_infer_affine();
    if (train) {
        DDPM_CHECK(s.training, "train-mode forward needs a training activation set");
        DDPM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 384 * (NUM_CONV + 1), stream));
    }
    const double count_local = (double)N;
    auto bn = [&](int l, bool pool) {
        // finalise batch statistics of layer l and apply BatchNorm+ReLU (train mode only)
        const ConvSpec& c = kConv[l];
        double m = count_local * c.hw * c.hw;
        // SyncBN: statistics of the GLOBAL batch.  With the peer-memory mailboxes the all-reduce happens inside
        // bn_finalize_kernel; without them (P2P unavailable) it is an NCCL call in front of it.
        int slot = -1;
        const double* src = gsum(l);
        if (sync_bn && comm && !(opt_dp_skip & 2)) {
            if (xr_ok && opt_bn_p2p) { slot = l - 1; src = lsum(l); }
            else allreduce_sums(lsum(l), gsum(l), 2 * c.cout);
        } else if (sync_bn && comm) {
            src = lsum(l);                 // timing-only mode: local statistics stand in for the global ones
        }
        if (sync_bn && comm) m *= world;
        bn_finalize_kernel<<<1, 256, 0, stream>>>(src, m, arr(c.bn + 1), arr(c.bn), arr(c.bn + 2), arr(c.bn + 3),
                                                  tr_mean[l], tr_istd[l], tr_scale[l], tr_shift[l], c.cout, 1e-5f, 0.1f,
                                                  update_running ? 1 : 0, xr, slot, gsum(l));
        long long work;
        if (pool) {
            work = (long long)N * 16 * 16 * (c.cout / 8);
            bn_apply_pool_kernel<TA><<<cdiv(work, 256), 256, 0, stream>>>(s.y[l].cview<TA>(), s.a[l].view<TA>(), s.p1.view<TA>(),
                                                                          s.y[l].g, s.p1.g, c.cout, tr_scale[l], tr_shift[l],
                                                                          (int)opt_train_reverse);
        } else {
            work = (long long)N * c.hw * c.hw * (c.cout / 8);
            bn_apply_kernel<TA><<<cdiv(work, 256), 256, 0, stream>>>(s.y[l].cview<TA>(), s.a[l].view<TA>(), s.y[l].g, c.cout,
                                                                     tr_scale[l], tr_shift[l], (int)opt_train_reverse);
        }
        DDPM_LAUNCH_CHECK();
        cnt_launches += 2;
    };
    auto layer = [&](int l, const Tensor& in0, const Tensor* in1) {
        NvtxRange r(train ? "fwd.train.L" : "fwd.infer.L", l);
        const ConvSpec& c = kConv[l];
        if (train) {
            conv3<TA, TG>(in0, in1, l, s.y[l], false, arr(c.b), 0, lsum(l));
            bn(l, l == 2);
        } else {
            // launch order of an evaluation: first conv, L2, pool, L3, L4, L5, L6, ConvTranspose, L7, L8, L9, L10; the tile
            // direction alternates along it (the first conv and the pool run front to back), see TcParams::rev
            conv3<TA, TG>(in0, in1, l, s.a[l], true, inf_shift[l], 1, nullptr, (l == 2 || l == 3 || l == 5 || l == 8 || l == 10) ? 1 : 0);
        }
    };

    // ---- down1.conv1 (+ folded embedding)
    {
        NvtxRange r(train ? "fwd.train.L" : "fwd.infer.L", 1);
        const ConvSpec& c = kConv[1];
        long long pixels = (long long)N * HW;
        Tensor& o = train ? s.y[1] : s.a[1];
        bool done = false;
        if (!train && !ts_dev && opt_conv1_tc && use_tc())
            done = tc::conv1_shared_t<TA>(stream, x_dev, Wimg, Ecls + (long long)(t_fixed - 1) * 9 * 64, inf_scale[1],
                                          inf_shift[1], 1, o.pos0<TA>(), o.g, opt_conv1_tc >= 2);
        if (!done) {
            conv1_kernel<TA><<<cdiv(pixels, CONV1_PIX_PER_BLOCK), 256, 0, stream>>>(x_dev, ts_dev, t_fixed, Wimg, Ecls,
                                                              train ? nullptr : inf_scale[1], train ? arr(c.b) : inf_shift[1],
                                                              train ? 0 : 1, o.view<TA>(), o.g, train ? lsum(1) : nullptr);
            DDPM_LAUNCH_CHECK();
        }
        cnt_launches += 1;
        if (train) bn(1, false);
    }
    layer(2, s.a[1], nullptr);
    if (!train) {  // MaxPool((2,2)) of h1 (train mode: fused into the BatchNorm apply of layer 2)
        long long work = (long long)N * 16 * 16 * 8;
        bn_apply_pool_kernel<TA><<<cdiv(work, 256), 256, 0, stream>>>(s.a[2].cview<TA>(), s.a[2].view<TA>(), s.p1.view<TA>(),
                                                                      s.a[2].g, s.p1.g, 64, nullptr, nullptr);
        DDPM_LAUNCH_CHECK();
        cnt_launches += 1;
    }
    layer(3, s.p1, nullptr);
    layer(4, s.a[3], nullptr);
    layer(5, s.a[4], nullptr);
    layer(6, s.a[5], nullptr);
    // ---- ConvTranspose((2,2), 128=>64, stride=2): GEMM [pos16][128] x [128][4*64] + pixel shuffle
    {
        NvtxRange r("fwd.convT");
        const Geo& gi = s.a[6].g;
        const Geo& go = s.u.g;
        bool done = false;
        if (use_tc())
            done = tc::up2<TA>(stream, s.a[6].pos0<TA>(), (const TA*)Wt, s.u.pos0<TA>(), gi, go, arr(kUpB), train ? 0 : 1);
        if (!done) {
            EpiUp2<TA> epi{s.u.view<TA>(), gi, go, arr(kUpB), 64, nullptr};
            launch_igemm_simt<TA, TA>(stream, s.a[6].cview<TA>(), 128, View<const TA>{nullptr, 0}, 0, (const TA*)Wt, 256, 1,
                                      gi.npos, MapId{gi.npos}, epi);
        }
        cnt_launches += 1;
    }
    layer(7, s.u, nullptr);
    layer(8, s.a[7], nullptr);
    layer(9, s.a[8], &s.a[2]);  // cat(up_h3, h1; dims=3): upsampled first, skip second (train_brain.jl:175)
    if (!skip_last) layer(10, s.a[9], nullptr);   // the sampler fuses layer 10 with the final conv + reverse update
}

inline void Engine::forward(ActSet& s, const float* x_dev, const int* ts_dev, int t_fixed, Mode mode, bool update_running) {
    DDPM_DISPATCH(prec, (forward_t<TA, TG>(s, x_dev, ts_dev, t_fixed, mode, update_running)));
}

template <typename TA, typename TG>
void Engine::final_conv_t(ActSet& s, float* eps_hat_dev) {
    long long work = (long long)s.N * HW * 8;
    final_conv_kernel<TA><<<cdiv(work, 256), 256, 0, stream>>>(s.a[10].cview<TA>(), s.a[10].g, arr(kFinalW), arr(kFinalB),
                                                               eps_hat_dev, 0, nullptr, nullptr, make_float4(0, 0, 0, 0), nullptr,
                                                               0u, 0);
    DDPM_LAUNCH_CHECK();
    cnt_launches += 1;
}

// Map every rank's mailbox into this process (cudaIpc handles exchanged through the NCCL communicator that was just
// created).  Any failure leaves xr_ok = false with the reason in xr_note: the NCCL all-reduce path is used instead.
inline void Engine::init_peer_mailboxes() {
    xr_ok = false;
    if (world < 2 || world > XR_MAX_WORLD) { xr_note = "world size outside 2..16"; return; }
    if (!nccl().AllGather) { xr_note = "ncclAllGather not found"; return; }
    try {
        DDPM_CUDA(cudaMalloc(&xr_own, XReduce::bytes()));
        DDPM_CUDA(cudaMemset(xr_own, 0, XReduce::bytes()));
        DDPM_CUDA(cudaMalloc(&xr.epoch, XR_SLOTS * sizeof(unsigned)));
        DDPM_CUDA(cudaMemset(xr.epoch, 0, XR_SLOTS * sizeof(unsigned)));
        cudaIpcMemHandle_t mine;
        DDPM_CUDA(cudaIpcGetMemHandle(&mine, xr_own));
        cudaIpcMemHandle_t* d_all = nullptr;
        DDPM_CUDA(cudaMalloc(&d_all, sizeof(cudaIpcMemHandle_t) * (size_t)(world + 1)));
        DDPM_CUDA(cudaMemcpy(d_all + world, &mine, sizeof mine, cudaMemcpyHostToDevice));
        DDPM_CUDA(cudaDeviceSynchronize());
        nccl().check(nccl().AllGather(d_all + world, d_all, sizeof(cudaIpcMemHandle_t), ncclChar, comm, stream),
                     "ncclAllGather(ipc handles)");
        DDPM_CUDA(cudaStreamSynchronize(stream));
        std::vector<cudaIpcMemHandle_t> all(world);
        DDPM_CUDA(cudaMemcpy(all.data(), d_all, sizeof(cudaIpcMemHandle_t) * (size_t)world, cudaMemcpyDeviceToHost));
        cudaFree(d_all);
        bool ok = true;
        for (int r = 0; r < world; ++r) {
            if (r == rank) { xr.peer[r] = xr_own; continue; }
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                xr_note = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e);
                ok = false;
                break;
            }
            xr_mapped[r] = p;
            xr.peer[r] = reinterpret_cast<double*>(p);
        }
        // every rank must take the same path: agree through a tiny all-reduce (min over ranks of `ok`)
        int* d_ok = nullptr;
        DDPM_CUDA(cudaMalloc(&d_ok, sizeof(int)));
        int h_ok = ok ? 1 : 0;
        DDPM_CUDA(cudaMemcpy(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice));
        nccl().check(nccl().AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, comm, stream), "ncclAllReduce(mailbox agreement)");
        DDPM_CUDA(cudaStreamSynchronize(stream));
        DDPM_CUDA(cudaMemcpy(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
        cudaFree(d_ok);
        xr.rank = rank; xr.world = world;
        xr_ok = h_ok == 1;
        if (xr_ok) xr_note = "peer mailboxes mapped";
        else if (ok) xr_note = "a peer could not map the mailboxes";
    } catch (const std::exception& ex) {
        xr_note = ex.what();
        xr_ok = false;
    }
}

inline void Engine::allreduce_sums(double* local, double* global, int n) {
    nccl().check(nccl().AllReduce(local, global, n, ncclFloat64, ncclSum, comm, stream), "ncclAllReduce(bn sums)");
}

// ------------------------------------------------------------------------------------ backward
template <typename TA, typename TG>
void Engine::backward_t(ActSet& s, const float* xt_dev, const int* ts_dev, const float* deps_dev, float alpha) {
    const int N = s.N;
    DDPM_CUDA(cudaMemsetAsync(G, 0, n_params * sizeof(float), stream));
    DDPM_CUDA(cudaMemsetAsync(misc_sums + 8, 0, sizeof(double) * 248, stream));
    const double count_local = (double)N;

    // BatchNorm(relu) backward of layer l: da (view) -> dy tensor
    // have_sums: the kernel that produced da already accumulated sum g, sum g*xhat into lsum(l) (fused epilogue)
    auto bn_bwd = [&](int l, View<const TG> da, Tensor& dy, bool have_sums) {
        NvtxRange r("bwd.bn.L", l);
        const ConvSpec& c = kConv[l];
        const Geo& g = s.y[l].g;
        long long pixels = (long long)N * c.hw * c.hw;
        int blocks = stride_blocks(pixels, BNB_PIX_PER_BLOCK, tc::state().num_sms, (int)opt_bnbwd_blocks);
        double m = count_local * c.hw * c.hw;
        if (!have_sums) {
            bn_reduce_linear_kernel<TA, TG, 1><<<lin_reduce_blocks(g.npos, tc::state().num_sms), 256, 0, stream>>>(
                s.y[l].cview<TA>(), da, g.npos, c.cout, tr_scale[l], tr_shift[l], tr_mean[l], tr_istd[l], lsum(l));
            cnt_launches += 1;
        }
        int slot = -1;
        const double* gs = gsum(l);
        if (sync_bn && comm && !(opt_dp_skip & 2)) {
            if (xr_ok && opt_bn_p2p) slot = NUM_CONV + l - 1;
            else allreduce_sums(lsum(l), gsum(l), 2 * c.cout);
        } else if (sync_bn && comm) {
            gs = lsum(l);
        }
        if (sync_bn && comm) m *= world;
        bn_bwd_means_kernel<<<1, 256, 0, stream>>>(lsum(l), gs, m, c.cout, bw_mg[l], bw_mgx[l], garr(c.bn), garr(c.bn + 1),
                                                   alpha, xr, slot, gsum(l));
        bn_bwd_kernel<TA, TG, 2><<<blocks, 256, 0, stream>>>(s.y[l].cview<TA>(), da, dy.view<TG>(), g, c.cout, tr_scale[l],
                                                             tr_shift[l], tr_mean[l], tr_istd[l], bw_mg[l], bw_mgx[l], lsum(l),
                                                             (int)opt_train_reverse);
        f64_to_f32_kernel<<<1, 128, 0, stream>>>(lsum(l) + 2 * c.cout, garr(c.b), c.cout, (double)alpha);
        DDPM_LAUNCH_CHECK();
        cnt_launches += 3;
    };
    // Data parallel: the gradient arena is all-reduced in five layer-ordered buckets (arena order = constructor order:
    // down1 | down2 | mid | up2 | up1+final); a bucket goes to the communication stream as soon as the backward pass
    // has written its last gradient, so only the last one (down1, 0.45 MB) is exposed.  Arrays [a0, a1).
    int n_bucket = 0;
    auto grad_bucket = [&](int a0, int a1) {
        if (!comm || (opt_dp_skip & 1)) return;
        if (capturing) {                       // recording: close the segment here, the collective is issued at replay
            seg_end();
            capturing->buckets.emplace_back(a0, a1);
            seg_begin();
            return;
        }
        issue_bucket(a0, a1, n_bucket++);
    };
    // weight gradient of a 3x3 conv: dW[co][tap][ci] = sum_p dy[p][co] * x[p + shift(tap)][ci]
    auto wgrad = [&](int l, const Tensor& dy, const Tensor& x, int ci_off) {
        NvtxRange r("bwd.wgrad.L", l);
        const ConvSpec& c = kConv[l];
        const Geo& g = dy.g;
        bool done = false;
        if (use_tc())
            done = tc::wgrad3x3<TG, TA>(stream, wg, dy.pos0<TG>(), c.cout, x.pos0<TA>(), x.C, g, garr(c.w), c.cin, ci_off, alpha);
        if (!done) {
            MapConv3 mapB{g.Wp, -(long long)g.guard, g.npos + g.guard};
            launch_wgrad_simt<TG, TA>(stream, dy.cview<TG>(), x.cview<TA>(), g.npos, 9, c.cout, x.C, MapId{g.npos}, mapB,
                                      IdxConv3{c.cin, ci_off}, alpha, garr(c.w));
        }
        cnt_launches += 1;
    };

    // memset the sums used by the backward (forward sums are no longer needed)
    DDPM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 384 * (NUM_CONV + 1), stream));

    // f<l>: the kernel that produced d(a_l) also accumulated the first BatchNorm-backward pass of layer l.  Layers 10 and 2
    // get their input gradient from HBM-bound elementwise kernels (final conv backward, pool/skip merge): fusing the
    // reduction there made those kernels latency-bound (measured 122 -> 275 us and 156 -> 376 us at B = 2048), so they
    // keep the streaming reduction kernel.
    const bool fz = use_tc() && opt_fuse_bn && sizeof(TA) == 2;
    bool f10 = false, f9 = false, f8 = false, f7 = false, f6 = false, f5 = false, f4 = false, f3 = false, f2 = false, f1 = false;
    // ---- final 1x1 conv
    {
        long long pixels = (long long)N * HW;
        final_bwd_kernel<TA, TG><<<stride_blocks(pixels, FINAL_BWD_PIX_PER_BLOCK, tc::state().num_sms, 4), 256, 0, stream>>>(
            s.a[10].cview<TA>(), s.g32a.view<TG>(), s.a[10].g, arr(kFinalW), deps_dev, misc_sums + 8);
        f64_to_f32_kernel<<<1, 128, 0, stream>>>(misc_sums + 8, garr(kFinalW), 64, (double)alpha);
        f64_to_f32_kernel<<<1, 32, 0, stream>>>(misc_sums + 72, garr(kFinalB), 1, (double)alpha);
        DDPM_LAUNCH_CHECK();
        cnt_launches += 3;
    }
    // ---- up1
    bn_bwd(10, s.g32a.cview<TG>(), s.g32b, f10);
    wgrad(10, s.g32b, s.a[9], 0);
    dgrad3<TA, TG>(s, s.g32b, 10, s.g32a, 64, 9, &f9);
    bn_bwd(9, s.g32a.cview<TG>(), s.g32b, f9);
    wgrad(9, s.g32b, s.a[8], 0);
    wgrad(9, s.g32b, s.a[2], 64);
    dgrad3<TA, TG>(s, s.g32b, 9, s.gcat, 128, 8, &f8);      // d(cat): channels 0..63 = d(a8), 64..127 = skip half of d(h1)
    grad_bucket(50, NUM_ARRAYS);   // up1 + final complete
    // ---- up2
    bn_bwd(8, s.gcat.cview<TG>(0), s.g32b, f8);
    wgrad(8, s.g32b, s.a[7], 0);
    dgrad3<TA, TG>(s, s.g32b, 8, s.g32a, 64, 7, &f7);
    bn_bwd(7, s.g32a.cview<TG>(), s.g32b, f7);
    wgrad(7, s.g32b, s.u, 0);
    dgrad3<TA, TG>(s, s.g32b, 7, s.g32a, 64);  // g32a = d(u)
    {   // ConvTranspose backward
        const Geo& gi = s.a[6].g;
        const Geo& go = s.u.g;
        long long pixels = (long long)N * HW;
        channel_sum_kernel<TG><<<stride_blocks(pixels, BNB_PIX_PER_BLOCK, tc::state().num_sms, 4), 256, 0, stream>>>(s.g32a.cview<TG>(), go, 64, misc_sums + 128);
        f64_to_f32_kernel<<<1, 128, 0, stream>>>(misc_sums + 128, garr(kUpB), 64, (double)alpha);
        // da6[in][ci] = sum_{q,co} du[outpos(in,q)][co] * Wtd[ci][q*64+co]
        bool done = false;
        if (use_tc()) {
            long long work = (long long)N * 16 * 16 * 4 * 8;
            unshuffle2_kernel<TG><<<cdiv(work, 256), 256, 0, stream>>>(s.g32a.cview<TG>(), s.gdu4.view<TG>(), go, gi, 64);
            tc::BnFuse bf{};
            if (fz) bf = bn_bwd_fuse(s, 6);
            done = tc::gemm_rows<TG>(stream, s.gdu4.pos0<TG>(), 256, (const TG*)Wtd, 128, s.g16a.pos0<TG>(), gi, &bf, &f6);
            cnt_launches += 1;
        }
        if (!done) {
            EpiConv<TG> epi{s.g16a.view<TG>(), gi, nullptr, nullptr, 0, nullptr};
            launch_igemm_simt<TG, TG>(stream, s.g32a.cview<TG>(), 64, View<const TG>{nullptr, 0}, 0, (const TG*)Wtd, 128, 4,
                                      gi.npos, MapUp2{gi, go}, epi);
        }
        // dW[a,b,co,ci] = sum_in du[outpos(in,q)][co] * a6[in][ci]
        bool wdone = false;
        if (done) {
            if constexpr (std::is_same<TG, TA>::value)
                wdone = tc::wgrad_up2<TG>(stream, wg, s.gdu4.pos0<TG>(), s.a[6].pos0<TA>(), gi, garr(kUpW), alpha);
        }
        if (!wdone)
            launch_wgrad_simt<TG, TA>(stream, s.g32a.cview<TG>(), s.a[6].cview<TA>(), gi.npos, 4, 64, 128, MapUp2{gi, go},
                                      MapValid{gi}, IdxUp2{64}, alpha, garr(kUpW));
        DDPM_LAUNCH_CHECK();
        cnt_launches += 4;
    }
    grad_bucket(kUpW, 50);         // up2 complete
    // ---- mid, down2
    bn_bwd(6, s.g16a.cview<TG>(), s.g16b, f6);
    wgrad(6, s.g16b, s.a[5], 0);
    dgrad3<TA, TG>(s, s.g16b, 6, s.g16a, 128, 5, &f5);
    bn_bwd(5, s.g16a.cview<TG>(), s.g16b, f5);
    wgrad(5, s.g16b, s.a[4], 0);
    dgrad3<TA, TG>(s, s.g16b, 5, s.g16a, 128, 4, &f4);
    grad_bucket(24, 36);           // mid complete
    bn_bwd(4, s.g16a.cview<TG>(), s.g16b, f4);
    wgrad(4, s.g16b, s.a[3], 0);
    dgrad3<TA, TG>(s, s.g16b, 4, s.g16a, 128, 3, &f3);
    bn_bwd(3, s.g16a.cview<TG>(), s.g16b, f3);
    wgrad(3, s.g16b, s.p1, 0);
    dgrad3<TA, TG>(s, s.g16b, 3, s.gp1, 64);
    grad_bucket(12, 24);           // down2 complete
    // ---- down1: h1 receives the skip half of d(cat) plus the MaxPool-routed gradient
    {
        long long work = (long long)N * 16 * 16 * 8;
        pool_bwd_merge_kernel<TA, TG><<<cdiv(work, 256), 256, 0, stream>>>(s.a[2].cview<TA>(), s.gcat.cview<TG>(64),
                                                                           s.gp1.cview<TG>(), s.g32a.view<TG>(), s.a[2].g,
                                                                           s.p1.g, 64);
        DDPM_LAUNCH_CHECK();
        cnt_launches += 1;
    }
    bn_bwd(2, s.g32a.cview<TG>(), s.g32b, f2);
    wgrad(2, s.g32b, s.a[1], 0);
    dgrad3<TA, TG>(s, s.g32b, 2, s.g32a, 64, 1, &f1);
    bn_bwd(1, s.g32a.cview<TG>(), s.g32b, f1);
    {   // first conv: image channel + folded embedding channels
        {
            // dy of the first conv -> per-image border-class sums (embedding channels) + image-channel weight gradient
            DDPM_CUDA(cudaMemsetAsync(l1_acc, 0, 576 * sizeof(double), stream));
            const int dt = std::is_same<TG, float>::value ? 0 : (std::is_same<TG, __half>::value ? 1 : 2);
            View<const __half> vh{reinterpret_cast<const __half*>(s.g32b.pos0<TG>()), s.g32b.C};
            View<const __nv_bfloat16> vb{reinterpret_cast<const __nv_bfloat16*>(s.g32b.pos0<TG>()), s.g32b.C};
            View<const float> vf{reinterpret_cast<const float*>(s.g32b.pos0<TG>()), s.g32b.C};
            if constexpr (sizeof(TG) == 2) {
                // 16-bit gradients: dy staged through shared memory by bulk asynchronous copies (l1_bwd.cuh)
                launch_l1_bwd_bulk<TG>(stream, s.g32b.pos0<TG>(), s.g32b.g, xt_dev, s.Ccls.as<float>(), l1_acc);
            } else {
                const int blocks = std::min(N, 2 * tc::state().num_sms);
                l1_bwd_fused_kernel<<<blocks, 256, 0, stream>>>(vh, vb, vf, dt, s.g32b.g, xt_dev, s.Ccls.as<float>(), l1_acc);
            }
            l1_wimg_finish_kernel<<<3, 256, 0, stream>>>(l1_acc, alpha, 129, garr(0));
        }
        l1_tap_sums_kernel<<<cdiv((long long)N * 576, 256), 256, 0, stream>>>(s.Ccls.as<float>(), s.S.as<float>(), N);
        View<const float> Sv{s.S.as<float>(), 576}, pev{d_pe, D};
        launch_wgrad_simt<float, float>(stream, Sv, pev, (long long)N, 1, 576, D, MapId{(long long)N}, MapTs{ts_dev, (long long)N},
                                        IdxEmb{129, 64}, alpha, garr(0));
        DDPM_LAUNCH_CHECK();
        cnt_launches += 4;
    }
    grad_bucket(0, 12);            // down1 complete: the only bucket whose all-reduce cannot hide behind backward work
    if (comm && !(opt_dp_skip & 1) && !capturing) join_comm();      // (replay joins after its last bucket)
}

// ------------------------------------------------------------------------------------ one training iteration
// Enqueues q_sample .. Adam .. weight re-packing on `stream`.  Inputs are resident in the set: s.x0 (or dataset rows
// gathered through s.idx), s.ts, s.eps -- host-supplied, or drawn here from Philox keyed by d_trng = [seed, first
// global image index, step].  Nothing in here depends on a host-side per-step scalar, so the sequence can be captured.
inline void Engine::train_enqueue(ActSet& s, bool gather, bool device_draws, bool update) {
    NvtxRange r("train_step");
    const int B = s.N;
    const float* x0 = gather ? d_dataset.as<float>() : s.x0.as<float>();
    const int* idx = gather ? s.idx.as<int>() : nullptr;
    long long n4 = (long long)B * HW / 4;
    if (device_draws) {
        randint_ts_dev_kernel<<<cdiv(B, 256), 256, 0, stream>>>(s.ts.as<int>(), B, T, d_trng);
        randn_train_dev_kernel<<<cdiv(n4, 256), 256, 0, stream>>>(s.eps.as<float>(), B, HW, d_trng);
        cnt_launches += 2;
    }
    qsample_kernel<<<cdiv(n4, 256), 256, 0, stream>>>(x0, idx, s.eps.as<float>(), s.ts.as<int>(), d_sqrt_ac, d_sqrt_1mac,
                                                      s.xt.as<float>(), B, HW);
    DDPM_LAUNCH_CHECK();
    cnt_launches += 1;
    forward(s, s.xt.as<float>(), s.ts.as<int>(), 0, Mode::Train, update);
    DDPM_DISPATCH(prec, (final_conv_t<TA, TG>(s, s.eps_hat.as<float>())));
    DDPM_CUDA(cudaMemsetAsync(misc_sums, 0, sizeof(double), stream));
    // loss = mean over the GLOBAL batch; every rank contributes its local sum / (B*world*HW)
    float inv_count = 1.f / ((float)B * (float)world * (float)HW);
    const float gs = grad_scale(B);
    mse_kernel<<<cdiv(n4, 256), 256, 0, stream>>>(s.eps_hat.as<float>(), s.eps.as<float>(), n4, inv_count * gs, misc_sums,
                                                  s.deps.as<float>());
    DDPM_LAUNCH_CHECK();
    cnt_launches += 1;
    DDPM_DISPATCH(prec, (backward_t<TA, TG>(s, s.xt.as<float>(), s.ts.as<int>(), s.deps.as<float>(), 1.f / gs)));
    if (update) {
        NvtxRange ra("adam");
        // overflow guard of the 16-bit gradient tensors: a non-finite value anywhere in the (all-reduced) gradient
        // skips the update (weights and moments untouched, beta^t not advanced) and bumps a counter
        grad_check_kernel<<<cdiv(n_params, 1024), 256, 0, stream>>>(G, n_params, d_tstate);
        adam_kernel<<<cdiv(n_params, 256), 256, 0, stream>>>(P, G, M1, M2, n_params, eta, b1, b2, aeps, d_tstate);
        adam_advance_kernel<<<1, 32, 0, stream>>>(d_tstate, b1, b2);
        DDPM_LAUNCH_CHECK();
        cnt_launches += 3;
        pack_weights();
    }
}

inline void Engine::train_core(ActSet& s, bool gather, bool device_draws, bool update, float* loss_out_host) {
    const int B = s.N;
    const int key = (gather ? 4 : 0) + (device_draws ? 2 : 0) + (update ? 1 : 0);
    bool replayed = false;
    // SyncBN through NCCL (mailboxes unavailable or switched off) would put collectives inside the segments: run eagerly
    const bool graph_ok = opt_train_graph && !(comm && sync_bn && !(xr_ok && opt_bn_p2p) && !(opt_dp_skip & 2));
    if (graph_ok) {
        auto it = s.train_graphs.find(key);
        if (it == s.train_graphs.end() && s.train_calls[key] >= 1) {
            // second call of this kind on this set: record it (the first, eager call initialised every lazily
            // created resource -- function attributes, occupancy queries, derived tables, NCCL channels)
            ecls_valid = false;                       // the recorded step must contain the embedding-fold refresh
            DDPM_CUDA(cudaStreamSynchronize(stream));
            DDPM_CUDA(cudaStreamSynchronize(comm_stream));
            TrainGraph tg;
            const long long before = cnt_launches;
            capturing = &tg;
            try {
                seg_begin();
                train_enqueue(s, gather, device_draws, update);
                seg_end();
            } catch (...) {
                cudaGraph_t g = nullptr;
                cudaStreamEndCapture(stream, &g);
                if (g) cudaGraphDestroy(g);
                capturing = nullptr;
                for (auto& ge : tg.segs) ActSet::destroy(ge);
                throw;
            }
            capturing = nullptr;
            tg.launches = cnt_launches - before;
            cnt_launches = before;
            it = s.train_graphs.emplace(key, tg).first;
        }
        if (it != s.train_graphs.end()) {
            TrainGraph& tg = it->second;
            for (size_t i = 0; i < tg.segs.size(); ++i) {
                DDPM_CUDA(cudaGraphLaunch(tg.segs[i].exec, stream));
                if (i < tg.buckets.size()) {
                    issue_bucket(tg.buckets[i].first, tg.buckets[i].second, (int)i);
                    if (i + 1 == tg.buckets.size()) join_comm();
                }
            }
            cnt_launches += tg.launches;
            replayed = true;
        }
    }
    if (!replayed) {
        if (update) ecls_valid = false;
        train_enqueue(s, gather, device_draws, update);
    }
    s.train_calls[key] += 1;
    if (update) { ecls_valid = false; infer_affine_valid = false; }    // the weights moved (also when replayed)
    if (loss_out_host) {
        double ls = 0;
        if (comm) {
            // global loss = sum of local sums / global count
            nccl().check(nccl().AllReduce(misc_sums, misc_sums, 1, ncclFloat64, ncclSum, comm, stream), "ncclAllReduce(loss)");
        }
        DDPM_CUDA(cudaMemcpyAsync(&ls, misc_sums, sizeof(double), cudaMemcpyDeviceToHost, stream));
        DDPM_CUDA(cudaStreamSynchronize(stream));
        const float inv_count = 1.f / ((float)B * (float)world * (float)HW);
        *loss_out_host = (float)(ls * (double)inv_count);   // mse_kernel accumulates the UNscaled squared error
    }
}

// ------------------------------------------------------------------------------------ sampling
template <typename TA, typename TG>
void Engine::sample_steps_t(ActSet& s, float* x_dev, const float* z_dev, int N, int t_start) {
    for (int t = t_start, k = 0; t >= 2; --t, ++k) {
        NvtxRange r("reverse_step");
        const float* zstep = z_dev ? z_dev + (size_t)k * N * HW : nullptr;
        if (!zstep) {
            // fresh noise of this step, Philox keyed by (seed, global image index, t): one fully parallel
            // launch (4 draws per thread) instead of one divergent generator lane per 8 in the fused kernel
            long long quads = (long long)N * HW / 4;
            randn_dev_kernel<<<cdiv(quads, 256), 256, 0, stream>>>(s.zstep.as<float>(), N, HW, s.rng.as<unsigned long long>(),
                                                                   (uint32_t)t);
            cnt_launches += 1;
            zstep = s.zstep.as<float>();
        }
        const float* sc = &h_samp[(size_t)(t - 1) * 4];
        if (use_tc() && opt_fuse_final) {
            forward_t<TA, TG>(s, x_dev, nullptr, t, Mode::Infer, false, true);
            bool fused = tc::conv3x3_final<TA>(stream, s.a[9].pos0<TA>(), (const TA*)Wfi[10], s.a[9].g, inf_shift[10], x_dev, zstep,
                                              arr(kFinalW), arr(kFinalB), sc, t == 2 ? 1 : 0,
                                              1 /* layer 10 of the alternating tile direction, see forward_t */);
            if (fused) {
                cnt_launches += 1;
                continue;
            }
            layer10_fallback<TA, TG>(s);
        } else {
            forward_t<TA, TG>(s, x_dev, nullptr, t, Mode::Infer, false);
        }
        long long work = (long long)N * HW * 8;
        final_conv_kernel<TA><<<cdiv(work, 256), 256, 0, stream>>>(
            s.a[10].cview<TA>(), s.a[10].g, arr(kFinalW), arr(kFinalB), nullptr, 1, x_dev,
            zstep, make_float4(sc[0], sc[1], sc[2], sc[3]), s.rng.as<unsigned long long>(), (uint32_t)t, t == 2 ? 1 : 0);
        cnt_launches += 1;
    }
    DDPM_LAUNCH_CHECK();
}

// reverse loop for the N images resident in s.x (in place); z from s.z (host-supplied) or Philox
inline void Engine::sample_chunk(ActSet& s, bool host_z, unsigned long long seed, long long first_index, int t_start) {
    prepare_ecls();
    prepare_infer_affine();
    unsigned long long rng[2] = {seed, (unsigned long long)first_index};
    DDPM_CUDA(cudaMemcpyAsync(s.rng.p, rng, sizeof rng, cudaMemcpyHostToDevice, stream));
    if (t_start < 2) return;
    float* x_dev = s.x.as<float>();
    const float* z_dev = host_z ? s.z.as<float>() : nullptr;
    const int N = s.N;
    if (!opt_use_graph) {
        DDPM_DISPATCH(prec, (sample_steps_t<TA, TG>(s, x_dev, z_dev, N, t_start)));
        return;
    }
    auto key = std::make_pair(t_start, host_z ? 1 : 0);
    auto it = s.graphs.find(key);
    if (it == s.graphs.end()) {
        DDPM_CUDA(cudaStreamSynchronize(stream));
        GraphEntry ge;
        long long before = cnt_launches;
        DDPM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        try {
            DDPM_DISPATCH(prec, (sample_steps_t<TA, TG>(s, x_dev, z_dev, N, t_start)));
        } catch (...) {
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        DDPM_CUDA(cudaStreamEndCapture(stream, &ge.graph));
        DDPM_CUDA(cudaGraphInstantiate(&ge.exec, ge.graph, 0));
        ge.launches = cnt_launches - before;
        cnt_launches = before;
        it = s.graphs.emplace(key, ge).first;
    }
    DDPM_CUDA(cudaGraphLaunch(it->second.exec, stream));
    cnt_launches += it->second.launches;
}

}  // namespace ddpm
