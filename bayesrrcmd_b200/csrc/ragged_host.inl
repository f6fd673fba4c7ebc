// Ragged layout, host side of the ABI (stage 1 placeholder: entry points exist and fail loudly).
static int ragged_compute_num(brr_gpu_ctx *, uint32_t, double, double, double *) { return fail(BRR_GPU_ESTATE, "ragged layout: not implemented yet"); }
static int ragged_apply_delta(brr_gpu_ctx *, uint32_t, double, double, double *) { return fail(BRR_GPU_ESTATE, "ragged layout: not implemented yet"); }
static int launch_ragged(brr_gpu_ctx *) { return fail(BRR_GPU_ESTATE, "ragged layout: not implemented yet"); }
extern "C" int brr_gpu_upload_ragged(brr_gpu_ctx *, const uint64_t *, const int32_t *, const double *) { return fail(BRR_GPU_ESTATE, "ragged layout: not implemented yet"); }
extern "C" int brr_gpu_upload_ragged_from_bed(brr_gpu_ctx *, const uint8_t *, size_t) { return fail(BRR_GPU_ESTATE, "ragged layout: not implemented yet"); }
extern "C" int brr_gpu_get_ragged_stats(brr_gpu_ctx *, double *) { return fail(BRR_GPU_ESTATE, "ragged layout: not implemented yet"); }
