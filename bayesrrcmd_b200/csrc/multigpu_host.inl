// Multi-GPU exchange, host side of the ABI: one context per GPU (usually one process per GPU), individuals
// sharded over the ranks.  Each rank allocates an exchange buffer (per-source mailboxes + round flags) with
// cudaMalloc, exports it as a CUDA IPC handle, and maps every peer's buffer.  The sweep kernel then pushes
// its share of each block's integer counts into the peers' mailboxes over NVLink and merges what the peers
// pushed (sweep_dense.cuh, "cross-GPU exchange"); nothing on this path goes through the host or NCCL.
extern "C" int brr_gpu_ipc_export(brr_gpu_ctx *c, void *handle_out)
{
    CHECK_CTX(c);
    if (!handle_out) return fail(BRR_GPU_EINVAL, "null handle buffer");
    if (c->cfg.layout != BRR_GPU_LAYOUT_DENSE2BIT) return fail(BRR_GPU_ESTATE, "the multi-GPU exchange exists for the dense 2-bit layout only (this context's layout is %d)", c->cfg.layout);
    if (c->cfg.world < 2 || !c->xchg) return fail(BRR_GPU_ESTATE, "context was created with world=1: nothing to exchange");
    static_assert(sizeof(cudaIpcMemHandle_t) == BRR_GPU_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->xchg));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int brr_gpu_ipc_attach(brr_gpu_ctx *c, const void *handles_all_ranks)
{
    CHECK_CTX(c);
    if (!handles_all_ranks) return fail(BRR_GPU_EINVAL, "null handles");
    if (c->cfg.world < 2 || !c->xchg) return fail(BRR_GPU_ESTATE, "context was created with world=1: nothing to attach");
    if (c->attached) return fail(BRR_GPU_ESTATE, "already attached");
    if (c->cfg.layout != BRR_GPU_LAYOUT_DENSE2BIT) return fail(BRR_GPU_ESTATE, "the multi-GPU exchange is implemented for the dense layout");
    std::vector<int *> ptrs(c->cfg.world, nullptr);
    for (int r = 0; r < c->cfg.world; ++r) {
        if (r == c->cfg.rank) { ptrs[r] = c->xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles_all_ranks) + size_t(r) * BRR_GPU_IPC_HANDLE_BYTES, sizeof(h));
        void *q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(BRR_GPU_ECUDA, "cudaIpcOpenMemHandle for rank %d failed: %s (the GPUs need peer access: NVLink / NVSwitch)", r, cudaGetErrorString(e));
        c->peer_opened.push_back(q);
        ptrs[r] = static_cast<int *>(q);
    }
    CUDA_TRY(cudaMemcpy(c->peer_ptrs_dev, ptrs.data(), sizeof(int *) * c->cfg.world, cudaMemcpyHostToDevice));
    uint32_t e_min = std::min<uint32_t>(kXchgCtas, c->nCTA);
    for (int r = 0; r < c->cfg.world; ++r) {
        int n = 0;
        CUDA_TRY(cudaMemcpy(&n, ptrs[r] + xchg_total_ints(c->cfg.world) - 8, sizeof(int), cudaMemcpyDeviceToHost));
        if (n < 1) return fail(BRR_GPU_ESTATE, "rank %d has not created its context yet (attach after ALL ranks exported)", r);
        e_min = std::min<uint32_t>(e_min, (uint32_t)n);
    }
    c->xchg_ctas = e_min;
    c->attached = true;
    return 0;
}

// In-process variant (several contexts of ONE process, e.g. tests or a single-process multi-GPU host): the
// same wiring from raw pointers instead of IPC handles.
extern "C" int brr_gpu_attach_local(brr_gpu_ctx **ctxs, int n)
{
    if (!ctxs || n < 2) return fail(BRR_GPU_EINVAL, "need >= 2 contexts");
    for (int i = 0; i < n; ++i) {
        brr_gpu_ctx *c = ctxs[i];
        if (!c || c->cfg.world != n || c->cfg.rank != i || !c->xchg) return fail(BRR_GPU_EINVAL, "context %d: expected rank %d of world %d", i, i, n);
        if (c->attached) return fail(BRR_GPU_ESTATE, "context %d already attached", i);
    }
    std::vector<int *> ptrs(n);
    uint32_t e_min = kXchgCtas;
    for (int i = 0; i < n; ++i) { ptrs[i] = ctxs[i]->xchg; e_min = std::min(e_min, ctxs[i]->nCTA); }
    for (int i = 0; i < n; ++i) {
        CUDA_TRY(cudaSetDevice(ctxs[i]->cfg.device));
        for (int j = 0; j < n; ++j)
            if (j != i && ctxs[j]->cfg.device != ctxs[i]->cfg.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->cfg.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return fail(BRR_GPU_ECUDA, "peer access %d -> %d: %s", ctxs[i]->cfg.device, ctxs[j]->cfg.device, cudaGetErrorString(e));
            }
        CUDA_TRY(cudaMemcpy(ctxs[i]->peer_ptrs_dev, ptrs.data(), sizeof(int *) * n, cudaMemcpyHostToDevice));
        ctxs[i]->xchg_ctas = e_min;
        ctxs[i]->attached = true;
    }
    return 0;
}

extern "C" uint32_t brr_gpu_num_blocks(brr_gpu_ctx *c) { return c ? (c->n_order + c->B - 1) / c->B : 0; }
