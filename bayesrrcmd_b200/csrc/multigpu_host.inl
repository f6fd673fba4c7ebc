// Multi-GPU exchange, host side of the ABI (stage 1 placeholder: entry points exist and fail loudly).
extern "C" int brr_gpu_ipc_export(brr_gpu_ctx *, void *) { return fail(BRR_GPU_ESTATE, "multi-GPU exchange: not implemented yet"); }
extern "C" int brr_gpu_ipc_attach(brr_gpu_ctx *, const void *) { return fail(BRR_GPU_ESTATE, "multi-GPU exchange: not implemented yet"); }
extern "C" int brr_gpu_block_stats(brr_gpu_ctx *, uint32_t, void **, uint32_t *, void **, uint32_t *) { return fail(BRR_GPU_ESTATE, "multi-GPU exchange: not implemented yet"); }
extern "C" int brr_gpu_block_finish(brr_gpu_ctx *, uint32_t) { return fail(BRR_GPU_ESTATE, "multi-GPU exchange: not implemented yet"); }
extern "C" uint32_t brr_gpu_num_blocks(brr_gpu_ctx *c) { return c ? (c->n_order + c->B - 1) / c->B : 0; }
