// Ragged / eigen-sparse layouts, host side of the ABI: build the CTA-major 16-bit segment store, launch
// the sweep, marker-level hooks.

static bool is_sparse_layout(int layout) { return layout == BRR_GPU_LAYOUT_RAGGED || layout == BRR_GPU_LAYOUT_EIGEN; }

// lists: for marker j, ones = idx[off[3j] .. off[3j+1]), twos = [off[3j+1] .. off[3j+2]), missing = [off[3j+2] .. off[3j+3])
static int ragged_build_store(brr_gpu_ctx *c, const uint64_t *off, const int32_t *idx, const double *stats)
{
    const uint32_t M = c->cfg.n_markers, N = c->cfg.n_inds, nCTA = c->nCTA, SL = c->SL;
    // validate + per-CTA totals
    std::vector<uint64_t> cta_total(nCTA, 0);
    std::vector<uint32_t> counts(size_t(M) * 4);
    for (uint32_t j = 0; j < M; ++j) {
        for (int l = 0; l < 3; ++l) {
            const uint64_t a = off[3 * size_t(j) + l], b = off[3 * size_t(j) + l + 1];
            if (b < a) return fail(BRR_GPU_EINVAL, "offsets not monotone at marker %u", j);
            int64_t prev = -1;
            for (uint64_t i = a; i < b; ++i) {
                const int32_t v = idx[i];
                if (v < 0 || uint32_t(v) >= N) return fail(BRR_GPU_EINVAL, "index %d of marker %u outside 0..%u", v, j, N - 1);
                if (int64_t(v) <= prev) return fail(BRR_GPU_EINVAL, "list %d of marker %u is not strictly ascending", l, j);
                prev = v;
                ++cta_total[uint32_t(v) / SL];
            }
        }
        const uint64_t n1 = off[3 * size_t(j) + 1] - off[3 * size_t(j)], n2 = off[3 * size_t(j) + 2] - off[3 * size_t(j) + 1],
                       nm = off[3 * size_t(j) + 3] - off[3 * size_t(j) + 2];
        if (n1 + n2 + nm > N) return fail(BRR_GPU_EINVAL, "marker %u lists more entries than individuals", j);
        counts[4 * size_t(j)] = uint32_t(N - n1 - n2 - nm); counts[4 * size_t(j) + 1] = uint32_t(n1);
        counts[4 * size_t(j) + 2] = uint32_t(n2); counts[4 * size_t(j) + 3] = uint32_t(nm);
    }
    std::vector<uint64_t> base(nCTA, 0);
    uint64_t total = 0;
    for (uint32_t k = 0; k < nCTA; ++k) {
        base[k] = total;
        if (cta_total[k] >= (uint64_t(1) << 32)) return fail(BRR_GPU_EINVAL, "too many entries in one slice");
        total += cta_total[k];
        total = (total + 7) & ~uint64_t(7);      // keep every CTA region 16-byte aligned
    }
    std::vector<uint16_t> idx16(std::max<uint64_t>(total, 1));
    std::vector<uint32_t> seg(size_t(nCTA) * M * 4);
    std::vector<uint64_t> fill(nCTA, 0);
    for (uint32_t j = 0; j < M; ++j) {
        for (uint32_t k = 0; k < nCTA; ++k) seg[(size_t(k) * M + j) * 4 + 0] = uint32_t(fill[k]);
        for (int l = 0; l < 3; ++l) {
            for (uint64_t i = off[3 * size_t(j) + l]; i < off[3 * size_t(j) + l + 1]; ++i) {
                const uint32_t v = uint32_t(idx[i]), k = v / SL;
                idx16[base[k] + fill[k]++] = uint16_t(v - k * SL);
            }
            for (uint32_t k = 0; k < nCTA; ++k) seg[(size_t(k) * M + j) * 4 + l + 1] = uint32_t(fill[k]);
        }
    }
    RaggedStore &r = c->rag;
    for (void *p : {(void *)r.idx16, (void *)r.seg, (void *)r.cta_base, (void *)r.stats}) if (p) cudaFree(p);
    r.idx16 = nullptr; r.seg = nullptr; r.cta_base = nullptr; r.stats = nullptr;
    int rc;
    if ((rc = dev_alloc(&r.idx16, idx16.size()))) return rc;
    if ((rc = dev_alloc(&r.seg, seg.size()))) return rc;
    if ((rc = dev_alloc(&r.cta_base, nCTA))) return rc;
    if ((rc = dev_alloc(&r.stats, size_t(M) * 4))) return rc;
    CUDA_TRY(cudaMemcpy(r.idx16, idx16.data(), idx16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.seg, seg.data(), seg.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.cta_base, base.data(), nCTA * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->counts_local, counts.data(), counts.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    r.n_entries = total; r.idx_bytes = idx16.size() * sizeof(uint16_t);
    r.stats_given = stats != nullptr;
    if (stats) CUDA_TRY(cudaMemcpy(r.stats, stats, sizeof(double) * 4 * M, cudaMemcpyHostToDevice));
    c->uploaded = true; c->finalized = false;
    return 0;
}

extern "C" int brr_gpu_upload_ragged(brr_gpu_ctx *c, const uint64_t *offsets, const int32_t *idx, const double *stats)
{
    CHECK_CTX(c);
    if (!is_sparse_layout(c->cfg.layout)) return fail(BRR_GPU_ESTATE, "context layout is not ragged/eigen");
    if (!offsets || (!idx && offsets[3 * size_t(c->cfg.n_markers)] > 0)) return fail(BRR_GPU_EINVAL, "null lists");
    return ragged_build_store(c, offsets, idx, stats);
}

// RaggedSparseMarkerBuilder::processAllele (reference src/raggedsparsemarkerbuilder.cpp:25-41) over a PLINK payload
extern "C" int brr_gpu_upload_ragged_from_bed(brr_gpu_ctx *c, const uint8_t *bed, size_t col_stride)
{
    CHECK_CTX(c);
    if (!is_sparse_layout(c->cfg.layout)) return fail(BRR_GPU_ESTATE, "context layout is not ragged/eigen");
    if (!bed) return fail(BRR_GPU_EINVAL, "null bed");
    const uint32_t M = c->cfg.n_markers, N = c->cfg.n_inds;
    if (col_stride < (size_t(N) + 3) / 4) return fail(BRR_GPU_EINVAL, "col_stride too small");
    std::vector<uint64_t> off(3 * size_t(M) + 1, 0);
    std::vector<int32_t> idx;
    std::vector<int32_t> l1, l2, lm;
    for (uint32_t j = 0; j < M; ++j) {
        const uint8_t *col = bed + size_t(j) * col_stride;
        l1.clear(); l2.clear(); lm.clear();
        for (uint32_t i = 0; i < N; ++i) {
            const unsigned two = (col[i >> 2] >> (2 * (i & 3))) & 3u;      // 00->2, 10->1, 11->0, 01->missing
            if (two == 2u) l1.push_back(int32_t(i));
            else if (two == 0u) l2.push_back(int32_t(i));
            else if (two == 1u) lm.push_back(int32_t(i));
        }
        off[3 * size_t(j)] = idx.size();
        idx.insert(idx.end(), l1.begin(), l1.end()); off[3 * size_t(j) + 1] = idx.size();
        idx.insert(idx.end(), l2.begin(), l2.end()); off[3 * size_t(j) + 2] = idx.size();
        idx.insert(idx.end(), lm.begin(), lm.end());
    }
    off[3 * size_t(M)] = idx.size();
    return ragged_build_store(c, off.data(), idx.data(), nullptr);
}

// Synthetic sparse workload generated on the device (bench / scale tests): the genotypes of brr_gpu_synth_dense as index lists.
extern "C" int brr_gpu_synth_ragged(brr_gpu_ctx *c, uint64_t seed, double maf_lo, double maf_hi, int log_uniform, double p_missing,
                                    uint32_t ind_offset)
{
    CHECK_CTX(c);
    if (!is_sparse_layout(c->cfg.layout)) return fail(BRR_GPU_ESTATE, "context layout is not ragged/eigen");
    if (!(maf_lo > 0.0 && maf_hi <= 0.5 && maf_lo <= maf_hi) || p_missing < 0.0 || p_missing >= 1.0) return fail(BRR_GPU_EINVAL, "bad maf / missing parameters");
    const uint32_t M = c->cfg.n_markers, nCTA = c->nCTA, SL = c->SL;
    std::vector<uint32_t> thr(size_t(M) * 3);
    synth_thresholds(seed, maf_lo, maf_hi, log_uniform, p_missing, 0, M, thr.data());
    uint32_t *dthr = nullptr, *dcnt = nullptr;
    int rc;
    if ((rc = dev_alloc(&dthr, size_t(M) * 3))) return rc;
    if ((rc = dev_alloc(&dcnt, size_t(nCTA) * M * 3))) { cudaFree(dthr); return rc; }
    auto cleanup = [&]() { cudaFree(dthr); cudaFree(dcnt); };
    cudaError_t e = cudaMemcpyAsync(dthr, thr.data(), sizeof(uint32_t) * 3 * M, cudaMemcpyHostToDevice, c->stream);
    const dim3 grid(nCTA, std::min<uint32_t>((M + 7) / 8, 2048));
    if (e == cudaSuccess) {
        synth_ragged_count_kernel<<<grid, 256, 0, c->stream>>>(seed, dthr, M, c->cfg.n_inds, ind_offset, nCTA, SL, dcnt);
        ++c->launches;
        e = cudaGetLastError();
    }
    std::vector<uint32_t> cnt(size_t(nCTA) * M * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), dcnt, cnt.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cleanup(); return fail(BRR_GPU_ECUDA, "synthetic generation failed: %s", cudaGetErrorString(e)); }
    // per-slice prefix sums -> seg, per-marker class counts of this shard
    std::vector<uint32_t> seg(size_t(nCTA) * M * 4), counts(size_t(M) * 4, 0);
    std::vector<uint64_t> base(nCTA, 0);
    uint64_t total = 0;
    for (uint32_t k = 0; k < nCTA; ++k) {
        uint64_t fill = 0;
        for (uint32_t j = 0; j < M; ++j) {
            const uint32_t *n = &cnt[(size_t(k) * M + j) * 3];
            uint32_t *sg = &seg[(size_t(k) * M + j) * 4];
            sg[0] = uint32_t(fill); sg[1] = uint32_t(fill + n[0]); sg[2] = uint32_t(fill + n[0] + n[1]); sg[3] = uint32_t(fill + n[0] + n[1] + n[2]);
            fill += uint64_t(n[0]) + n[1] + n[2];
            counts[4 * size_t(j) + 1] += n[0]; counts[4 * size_t(j) + 2] += n[1]; counts[4 * size_t(j) + 3] += n[2];
        }
        if (fill >= (uint64_t(1) << 32)) { cleanup(); return fail(BRR_GPU_EINVAL, "too many entries in one slice"); }
        base[k] = total;
        total = (total + fill + 7) & ~uint64_t(7);
    }
    for (uint32_t j = 0; j < M; ++j)
        counts[4 * size_t(j)] = c->cfg.n_inds - counts[4 * size_t(j) + 1] - counts[4 * size_t(j) + 2] - counts[4 * size_t(j) + 3];
    RaggedStore &r = c->rag;
    for (void *q : {(void *)r.idx16, (void *)r.seg, (void *)r.cta_base, (void *)r.stats}) if (q) cudaFree(q);
    r.idx16 = nullptr; r.seg = nullptr; r.cta_base = nullptr; r.stats = nullptr;
    if ((rc = dev_alloc(&r.idx16, std::max<uint64_t>(total, 1))) || (rc = dev_alloc(&r.seg, seg.size())) || (rc = dev_alloc(&r.cta_base, nCTA)) ||
        (rc = dev_alloc(&r.stats, size_t(M) * 4))) { cleanup(); return rc; }
    e = cudaMemcpyAsync(r.seg, seg.data(), seg.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r.cta_base, base.data(), nCTA * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->counts_local, counts.data(), counts.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        synth_ragged_fill_kernel<<<grid, 256, 0, c->stream>>>(seed, dthr, M, c->cfg.n_inds, ind_offset, nCTA, SL, r.seg, r.cta_base, r.idx16);
        ++c->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) return fail(BRR_GPU_ECUDA, "synthetic generation failed: %s", cudaGetErrorString(e));
    r.n_entries = total; r.idx_bytes = std::max<uint64_t>(total, 1) * sizeof(uint16_t);
    r.stats_given = false;
    c->uploaded = true; c->finalized = false;
    return 0;
}

extern "C" int brr_gpu_get_ragged_stats(brr_gpu_ctx *c, double *stats)
{
    CHECK_CTX(c);
    if (!is_sparse_layout(c->cfg.layout) || !stats) return fail(BRR_GPU_EINVAL, "ragged/eigen layout and non-null output required");
    if (!c->finalized) return fail(BRR_GPU_ESTATE, "markers not finalized");
    CUDA_TRY(cudaMemcpy(stats, c->rag.stats, sizeof(double) * 4 * c->cfg.n_markers, cudaMemcpyDeviceToHost));
    return 0;
}

static int launch_ragged(brr_gpu_ctx *c)
{
    RaggedSweepParams p{};
    p.n_local = c->cfg.n_inds; p.n_total = c->cfg.n_inds_total; p.M = c->cfg.n_markers; p.K = c->cfg.n_components;
    p.nG = c->cfg.n_groups; p.W = c->async_tokens ? c->async_tokens : c->B; p.nCTA = c->nCTA; p.SL = c->SL;
    p.eigen = c->cfg.layout == BRR_GPU_LAYOUT_EIGEN;
    p.idx16 = c->rag.idx16; p.seg = c->rag.seg; p.cta_base = c->rag.cta_base; p.stats = c->rag.stats; p.group = c->group;
    p.order = c->order; p.n_order = c->n_order; p.rand_uz = c->rand_uz; p.pi = c->pi; p.sigmaG = c->sigmaG; p.cva = c->cva;
    p.sigmaE = c->sigmaE; p.eps_sum_in = c->eps_sum_in; p.async_mode = c->async_tokens ? 1u : 0u;
    p.eps = c->eps; p.beta = c->beta; p.comp = c->comp; p.acum = c->acum;
    p.part_dot = c->part_dot; p.bar = c->bar; p.cta_red = c->cta_red;
    p.v = c->v; p.betasq = c->betasq; p.out_scalars = c->out_scalars; p.out_counts = c->out_counts;
    CUDA_TRY(cudaMemsetAsync(c->bar, 0, sizeof(unsigned) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->out_counts, 0, sizeof(unsigned) * 4, c->stream));      // [2] = error flags read by sweep_fetch
    void *args[] = {&p};
    CUDA_TRY(cudaFuncSetAttribute(sweep_ragged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));   // per-function state: see launch_dense
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)sweep_ragged_kernel, dim3(c->nCTA), dim3(kThreads), args, c->smem_bytes, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->launches += 1; c->last_launches = 1;
    return 0;
}

static int ragged_fetch_stats(brr_gpu_ctx *c, uint32_t marker, double st[4])
{
    CUDA_TRY(cudaMemcpy(st, c->rag.stats + 4 * size_t(marker), sizeof(double) * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// SparseBayesRKernel::computeNum (reference src/sparsebayesrkernel.cpp:24-27)
static int ragged_compute_num(brr_gpu_ctx *c, uint32_t marker, double beta_old, double eps_sum, double *num_out)
{
    double st[4];
    int rc = ragged_fetch_stats(c, marker, st);
    if (rc) return rc;
    const int eigen = c->cfg.layout == BRR_GPU_LAYOUT_EIGEN;
    ragged_dot_kernel<<<c->nCTA, 256, 0, c->stream>>>(c->rag.idx16, c->rag.seg, c->rag.cta_base, c->cfg.n_markers, marker, eigen, st[0],
                                                      c->eps, c->SL, c->cta_red);
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
    std::vector<double> h(c->nCTA);
    CUDA_TRY(cudaMemcpyAsync(h.data(), c->cta_red, sizeof(double) * c->nCTA, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double s = 0.0;
    for (double x : h) s += x;
    const double NM1 = double(c->cfg.n_inds_total) - 1.0;
    *num_out = (c->cfg.rank == 0 ? beta_old * NM1 - st[0] * eps_sum / st[1] : 0.0) + s / st[1];
    return 0;
}

// Ragged/EigenBayesRKernel::calculateEpsilonChange + computeEpsilonSumUpdate
// (reference src/raggedbayesrkernel.cpp:10-30, src/eigenbayesrkernel.cpp:10-15, src/sparsebayesrkernel.cpp:29-34)
static int ragged_apply_delta(brr_gpu_ctx *c, uint32_t marker, double beta_old, double beta_new, double *eps_sum_delta)
{
    double st[4];
    int rc = ragged_fetch_stats(c, marker, st);
    if (rc) return rc;
    const double dB = beta_old - beta_new;
    ragged_update_kernel<<<c->nCTA, 256, 0, c->stream>>>(c->rag.idx16, c->rag.seg, c->rag.cta_base, c->cfg.n_markers, marker,
                                                         dB * st[0] / st[1], dB / st[1], c->eps, c->cfg.n_inds, c->SL);
    ++c->launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (eps_sum_delta) *eps_sum_delta = dB * st[3] / st[1] - dB * st[0] * double(c->cfg.n_inds_total) / st[1];
    return 0;
}
