// sfh_group.cuh -- single-process multi-GPU: ONE host process shards the stack by bin rows over N GPUs (included by sfh_api.cu).
//
// Every reference caller is a single Julia process (fit_sfh: fitting/hierarchical/generic_fitting.jl:242-411; fit_templates:
// fitting/solvers.jl:82-90), so a 40 GB stack must be shardable without adopting a process launcher.  A group owns one
// sfh_stack + sfh_ctx per GPU, enables peer access between all pairs, wires the per-rank inboxes of the one-shot NVLink exchange
// (the finalize kernel's last block) with plain peer pointers -- no CUDA-IPC, no NCCL -- and hands the caller ONE context, the
// primary, which every fused entry point accepts: sfh_eval_fg / sfh_eval_fg_hier / sfh_hier_bind on it fan out to all GPUs, so the
// native BFGS / L-BFGS-B drivers and the Julia binding run unchanged on a group.  GPU 0 is driven by the calling thread, GPUs
// 1..N-1 by persistent worker threads that spin briefly for the next evaluation (a fit issues one every ~1 ms) and sleep
// otherwise.  Paths that reduce with NCCL (two-pass, batched walkers, helpers) are refused on a group context.
#pragma once

#include <atomic>
#include <condition_variable>
#include <thread>

struct sfh_group {
    int ndev = 0;
    std::vector<int> devices;
    std::vector<sfh_stack *> stacks;
    std::vector<sfh_ctx *> ctxs;
    std::vector<std::thread> workers;
    // one evaluation at a time; the job is published by bumping `gen`
    std::mutex call_mu;
    struct Job {
        int kind = 0;   // 1 fg, 2 hierarchical fg, 3 bind, 9 exit
        const double *x = nullptr, *x2 = nullptr;
        int want_G = 0, mh_kind = 0, disp_kind = 0;
        double mh_fixed[4] = {0, 0, 0, 0};
        uint8_t free_mask[4] = {1, 1, 1, 0};
        bool has_mask = false;
    } job;
    std::atomic<uint64_t> gen{0};
    std::atomic<int> done{0};
    std::atomic<int> sleepers{0};
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> status;
    std::vector<std::string> errs;
};

// bin rows of shard i of n: contiguous, balanced, interior boundaries on multiples of `align` bins (whole panels of any tile width)
extern "C" int sfh_shard_rows(int64_t nbins, int nshards, int i, int64_t align, int64_t *row_begin, int64_t *row_end) {
    if (nbins < 0 || nshards < 1 || i < 0 || i >= nshards || align < 1 || !row_begin || !row_end) return fail(SFH_ERR_INVALID_ARG, "bad argument");
    int64_t per = (nbins + nshards - 1) / nshards;
    per = (per + align - 1) / align * align;
    *row_begin = std::min<int64_t>(nbins, (int64_t)i * per);
    *row_end = std::min<int64_t>(nbins, (int64_t)(i + 1) * per);
    return SFH_OK;
}

namespace {
int group_run_local(sfh_group *g, int i, const sfh_group::Job &j, double *neg_logL, double *G, double *composite_out, int64_t *n_ages) {
    sfh_ctx *c = g->ctxs[(size_t)i];
    switch (j.kind) {
    case 1: return eval_fg_local(c, j.x, neg_logL, G, composite_out, j.want_G);
    case 2: return eval_fg_hier_local(c, j.mh_kind, j.mh_fixed, j.disp_kind, j.x, j.has_mask ? j.free_mask : nullptr, neg_logL, G, j.want_G);
    case 3: return hier_bind_local(c, j.x, j.x2, n_ages);
    default: return SFH_OK;
    }
}

void group_worker(sfh_group *g, int i) {
    cudaSetDevice(g->devices[(size_t)i]);
    uint64_t seen = 0;
    for (;;) {
        // wait for the next job: spin (~100 us), then sleep
        int spins = 0;
        while (g->gen.load(std::memory_order_acquire) == seen) {
            if (++spins < 200000) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                continue;
            }
            std::unique_lock<std::mutex> lk(g->mu);
            g->sleepers.fetch_add(1);
            g->cv.wait(lk, [&] { return g->gen.load(std::memory_order_acquire) != seen; });
            g->sleepers.fetch_sub(1);
        }
        seen = g->gen.load(std::memory_order_acquire);
        const sfh_group::Job j = g->job;
        if (j.kind == 9) return;
        const int st = guarded([&]() -> int { return group_run_local(g, i, j, nullptr, nullptr, nullptr, nullptr); });
        g->status[(size_t)i] = st;
        if (st != SFH_OK) g->errs[(size_t)i] = g_err;
        g->done.fetch_add(1, std::memory_order_release);
    }
}

// publish `j`, run rank 0 on the calling thread, wait for the others
int group_dispatch(sfh_group *g, const sfh_group::Job &j, double *neg_logL, double *G, double *composite_out, int64_t *n_ages) {
    std::lock_guard<std::mutex> call(g->call_mu);
    g->job = j;
    g->done.store(0, std::memory_order_relaxed);
    g->gen.fetch_add(1, std::memory_order_release);
    if (g->sleepers.load() > 0) {
        std::lock_guard<std::mutex> lk(g->mu);
        g->cv.notify_all();
    }
    int st = group_run_local(g, 0, j, neg_logL, G, composite_out, n_ages);
    std::string err0 = st != SFH_OK ? g_err : std::string();
    int spins = 0;
    while (g->done.load(std::memory_order_acquire) < g->ndev - 1) {
        if (++spins > 100000) std::this_thread::yield();
    }
    for (int i = 1; i < g->ndev && st == SFH_OK; ++i)
        if (g->status[(size_t)i] != SFH_OK) { st = g->status[(size_t)i]; err0 = "GPU " + std::to_string(g->devices[(size_t)i]) + ": " + g->errs[(size_t)i]; }
    if (st != SFH_OK) g_err = err0;
    return st;
}

void group_free(sfh_group *g) {
    if (!g) return;
    if (!g->workers.empty()) {
        {
            std::lock_guard<std::mutex> call(g->call_mu);
            g->job = sfh_group::Job();
            g->job.kind = 9;
            g->gen.fetch_add(1, std::memory_order_release);
            std::lock_guard<std::mutex> lk(g->mu);
            g->cv.notify_all();
        }
        for (auto &t : g->workers) if (t.joinable()) t.join();
    }
    for (sfh_ctx *c : g->ctxs) if (c) { c->group = nullptr; sfh_ctx_destroy_impl(c); }
    for (sfh_stack *s : g->stacks) sfh_stack_destroy_impl(s);
    delete g;
}

// stacks exist: contexts, peer access, inboxes, workers
int group_finish(sfh_group *g) {
    const int n = g->ndev;
    for (int i = 0; i < n; ++i) {
        if (!g->stacks[(size_t)i]->fused)
            return fail(SFH_ERR_UNSUPPORTED, "shard %d (GPU %d, rows [%lld, %lld)) did not get a fused tiling: a group needs every shard on the fused path",
                        i, g->devices[(size_t)i], (long long)g->stacks[(size_t)i]->row_begin, (long long)g->stacks[(size_t)i]->row_end);
        sfh_ctx *c = nullptr;
        SFH_TRY(sfh_ctx_create_impl(g->stacks[(size_t)i], nullptr, &c));
        g->ctxs[(size_t)i] = c;
    }
    if (n > 1) {
        for (int i = 0; i < n; ++i) {
            CU_TRY(cudaSetDevice(g->devices[(size_t)i]));
            for (int k = 0; k < n; ++k) {
                if (k == i) continue;
                int can = 0;
                CU_TRY(cudaDeviceCanAccessPeer(&can, g->devices[(size_t)i], g->devices[(size_t)k]));
                if (!can) return fail(SFH_ERR_UNSUPPORTED, "GPU %d cannot access GPU %d's memory (no NVLink / PCIe peer path)", g->devices[(size_t)i], g->devices[(size_t)k]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[(size_t)k], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU_TRY(e);
                (void)cudaGetLastError();
            }
            SFH_TRY(p2p_alloc_inbox(g->ctxs[(size_t)i], n));
        }
        std::vector<double *> peers((size_t)n);
        for (int i = 0; i < n; ++i) peers[(size_t)i] = g->ctxs[(size_t)i]->d_inbox;
        for (int i = 0; i < n; ++i) {
            CU_TRY(cudaSetDevice(g->devices[(size_t)i]));
            SFH_TRY(p2p_attach(g->ctxs[(size_t)i], n, i, peers));
        }
    }
    for (int i = 0; i < n; ++i) { g->ctxs[(size_t)i]->group = g; g->ctxs[(size_t)i]->group_primary = (i == 0); }
    g->status.assign((size_t)n, SFH_OK);
    g->errs.assign((size_t)n, std::string());
    for (int i = 1; i < n; ++i) g->workers.emplace_back(group_worker, g, i);
    return SFH_OK;
}

int group_alloc(sfh_group **out, const int *devices, int ndev, sfh_group **gp) {
    if (!out) return fail(SFH_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { (void)cudaGetLastError(); return fail(SFH_ERR_NO_DEVICE, "no CUDA device"); }
    if (ndev < 1 || ndev > 32 || ndev > have) return fail(SFH_ERR_INVALID_ARG, "ndev = %d with %d visible GPU(s)", ndev, have);
    sfh_group *g = new (std::nothrow) sfh_group();
    if (!g) return fail(SFH_ERR_OOM, "host allocation failed");
    g->ndev = ndev;
    for (int i = 0; i < ndev; ++i) {
        const int d = devices ? devices[i] : i;
        if (d < 0 || d >= have || std::find(g->devices.begin(), g->devices.end(), d) != g->devices.end()) {
            delete g;
            return fail(SFH_ERR_INVALID_ARG, "bad or repeated device ordinal %d", d);
        }
        g->devices.push_back(d);
    }
    g->stacks.assign((size_t)ndev, nullptr);
    g->ctxs.assign((size_t)ndev, nullptr);
    *gp = g;
    return SFH_OK;
}

sfh_opts group_shard_opts(const sfh_opts *opts, int device, int64_t nbins, int ndev, int i) {
    sfh_opts o{};
    if (opts) o = *opts;
    o.struct_size = (int32_t)sizeof(sfh_opts);
    o.device = device;
    int64_t b = 0, e = 0;
    sfh_shard_rows(nbins, ndev, i, 128, &b, &e);
    o.row_begin = b; o.row_end = e;   // (callers reject empty shards: (0, 0) would mean "all rows" to sfh_stack_create)
    return o;
}
}  // namespace

extern "C" int sfh_group_create(sfh_group **out, const void *models, int64_t nbins, int64_t ntemplates, int dtype, const void *data,
                                int data_dtype, const int *devices, int ndev, const sfh_opts *opts) {
    return guarded([&]() -> int {
        if (opts && opts->struct_size != (int32_t)sizeof(sfh_opts)) return fail(SFH_ERR_INVALID_ARG, "sfh_opts.struct_size mismatch");
        if (opts && (opts->row_begin || opts->row_end)) return fail(SFH_ERR_INVALID_ARG, "a group shards the rows itself: leave row_begin/row_end 0");
        if (nbins < (int64_t)ndev * 128) return fail(SFH_ERR_SHAPE, "%lld bins are too few to shard over %d GPUs (>= 128 bins per GPU)", (long long)nbins, ndev);
        sfh_group *g = nullptr;
        SFH_TRY(group_alloc(out, devices, ndev, &g));
        int st = SFH_OK;
        for (int i = 0; i < ndev && st == SFH_OK; ++i) {
            const sfh_opts o = group_shard_opts(opts, g->devices[(size_t)i], nbins, ndev, i);
            if (o.row_end <= o.row_begin) { st = fail(SFH_ERR_SHAPE, "shard %d of %d would be empty", i, ndev); break; }
            st = sfh_stack_create_impl(&g->stacks[(size_t)i], models, nbins, ntemplates, dtype, data, data_dtype, &o);
        }
        if (st == SFH_OK) st = group_finish(g);
        if (st != SFH_OK) { const std::string keep = g_err; group_free(g); g_err = keep; return st; }
        *out = g;
        return SFH_OK;
    });
}

extern "C" int sfh_group_create_synthetic(sfh_group **out, int64_t nbins, int64_t ntemplates, int dtype, uint64_t seed, double scale,
                                          const double *x_true, const int *devices, int ndev, const sfh_opts *opts) {
    return guarded([&]() -> int {
        if (opts && opts->struct_size != (int32_t)sizeof(sfh_opts)) return fail(SFH_ERR_INVALID_ARG, "sfh_opts.struct_size mismatch");
        if (opts && (opts->row_begin || opts->row_end)) return fail(SFH_ERR_INVALID_ARG, "a group shards the rows itself: leave row_begin/row_end 0");
        if (nbins < (int64_t)ndev * 128) return fail(SFH_ERR_SHAPE, "%lld bins are too few to shard over %d GPUs (>= 128 bins per GPU)", (long long)nbins, ndev);
        sfh_group *g = nullptr;
        SFH_TRY(group_alloc(out, devices, ndev, &g));
        int st = SFH_OK;
        for (int i = 0; i < ndev && st == SFH_OK; ++i) {
            const sfh_opts o = group_shard_opts(opts, g->devices[(size_t)i], nbins, ndev, i);
            if (o.row_end <= o.row_begin) { st = fail(SFH_ERR_SHAPE, "shard %d of %d would be empty", i, ndev); break; }
            st = sfh_stack_create_synthetic_impl(&g->stacks[(size_t)i], nbins, ntemplates, dtype, seed, scale, x_true, &o);
        }
        if (st == SFH_OK) st = group_finish(g);
        if (st != SFH_OK) { const std::string keep = g_err; group_free(g); g_err = keep; return st; }
        *out = g;
        return SFH_OK;
    });
}

extern "C" int sfh_group_destroy(sfh_group *g) {
    return guarded([&]() -> int { group_free(g); return SFH_OK; });
}

extern "C" int sfh_group_ctx(sfh_group *g, sfh_ctx **primary) {
    return guarded([&]() -> int {
        if (!g || !primary) return fail(SFH_ERR_INVALID_ARG, "NULL argument");
        *primary = g->ctxs[0];
        return SFH_OK;
    });
}

extern "C" int sfh_group_info(const sfh_group *g, int *ndev, sfh_info *infos) {
    return guarded([&]() -> int {
        if (!g) return fail(SFH_ERR_INVALID_ARG, "NULL group");
        if (ndev) *ndev = g->ndev;
        if (infos)
            for (int i = 0; i < g->ndev; ++i) SFH_TRY(sfh_stack_info_impl(g->stacks[(size_t)i], &infos[i]));
        return SFH_OK;
    });
}

// mean device time per evaluation over `reps` back-to-back group evaluations, max over the GPUs (bench plumbing).  Coefficients
// are already resident after one ordinary evaluation; the loop is driven exactly like sfh_eval_fg drives it, minus the host copies.
extern "C" int sfh_group_time_fg(sfh_group *g, const double *coeffs, int reps, int want_G, double *ms_per_eval_out) {
    return guarded([&]() -> int {
        if (!g || !coeffs || reps < 1 || !ms_per_eval_out) return fail(SFH_ERR_INVALID_ARG, "bad argument");
        double nl = 0.0;
        SFH_TRY(sfh_eval_fg_impl(g->ctxs[0], coeffs, &nl, nullptr, nullptr));   // warm-up + coefficients resident on every GPU
        std::lock_guard<std::mutex> call(g->call_mu);   // workers idle: this thread drives every stream itself
        const int n = g->ndev;
        std::vector<cudaEvent_t> e0((size_t)n), e1((size_t)n);
        int st = SFH_OK;
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(g->devices[(size_t)i]);
            cudaEventCreate(&e0[(size_t)i]); cudaEventCreate(&e1[(size_t)i]);
        }
        for (int r = -2; r < reps && st == SFH_OK; ++r) {
            for (int i = 0; i < n && st == SFH_OK; ++i) {
                sfh_ctx *c = g->ctxs[(size_t)i];
                cudaSetDevice(g->devices[(size_t)i]);
                if (r == 0) cudaEventRecord(e0[(size_t)i], c->stream);   // after two untimed, all-reduced (hence rank-aligning) evaluations
                st = enqueue_fg_impl(c, c->d_coeffs, c->d_out, want_G, false);
            }
        }
        double worst = 0.0;
        for (int i = 0; i < n; ++i) {
            sfh_ctx *c = g->ctxs[(size_t)i];
            cudaSetDevice(g->devices[(size_t)i]);
            cudaEventRecord(e1[(size_t)i], c->stream);
            const cudaError_t se = cudaStreamSynchronize(c->stream);
            float ms = 0.f;
            if (se == cudaSuccess) cudaEventElapsedTime(&ms, e0[(size_t)i], e1[(size_t)i]);
            else if (st == SFH_OK) st = fail(SFH_ERR_CUDA, "sfh_group_time_fg: %s", cudaGetErrorString(se));
            worst = std::max(worst, (double)ms);
            cudaEventDestroy(e0[(size_t)i]); cudaEventDestroy(e1[(size_t)i]);
        }
        if (st == SFH_OK) *ms_per_eval_out = worst / reps;
        return st;
    });
}

static int group_eval_fg(sfh_group *g, const double *coeffs, double *neg_logL, double *G, double *composite_out) {
    if (composite_out) return fail(SFH_ERR_UNSUPPORTED, "composite_out is not gathered across a group's shards");
    sfh_group::Job j;
    j.kind = 1; j.x = coeffs; j.want_G = G != nullptr;
    return group_dispatch(g, j, neg_logL, G, nullptr, nullptr);
}

static int group_eval_fg_hier(sfh_group *g, int mh_kind, const double *mh_fixed, int disp_kind, const double *variables,
                              const uint8_t *free_mask, double *neg_logL, double *G) {
    if (!mh_fixed) return fail(SFH_ERR_INVALID_ARG, "mh_fixed is NULL");
    sfh_group::Job j;
    j.kind = 2; j.x = variables; j.want_G = G != nullptr; j.mh_kind = mh_kind; j.disp_kind = disp_kind;
    const int nfix = (mh_kind == SFH_MH_LOG_AMR) ? 4 : 1;
    for (int i = 0; i < nfix; ++i) j.mh_fixed[i] = mh_fixed[i];
    j.has_mask = free_mask != nullptr;
    if (free_mask) for (int i = 0; i < 3; ++i) j.free_mask[i] = free_mask[i];
    return group_dispatch(g, j, neg_logL, G, nullptr, nullptr);
}

static int group_hier_bind(sfh_group *g, const double *logAge, const double *MH, int64_t *n_ages_out) {
    sfh_group::Job j;
    j.kind = 3; j.x = logAge; j.x2 = MH;
    return group_dispatch(g, j, nullptr, nullptr, nullptr, n_ages_out);
}

// one-off helper (renormalize_x0, mdf_amr): the shards' column sums are added on the host in rank order -- no collective needed
static int group_column_sums(sfh_group *g, double *colsums_out) {
    std::lock_guard<std::mutex> call(g->call_mu);
    const int64_t nt = g->stacks[0]->nt;
    std::vector<double> part((size_t)std::max<int64_t>(nt, 1));
    for (int64_t j = 0; j < nt; ++j) colsums_out[j] = 0.0;
    for (int i = 0; i < g->ndev; ++i) {
        SFH_TRY(column_sums_local(g->ctxs[(size_t)i], part.data(), false));
        for (int64_t j = 0; j < nt; ++j) colsums_out[j] += part[(size_t)j];
    }
    return SFH_OK;
}
