// Host side of the STFT-domain kernels (stft.cuh): per-device tables, workspace carving, the Griffin-Lim iteration loop.
// Included at the end of zs_ae.cu (one translation unit, one libzsae.so).
#pragma once
#include <math.h>

#include "stft.cuh"

struct StftTables { float2* w1024 = nullptr; float* win = nullptr; };
static StftTables g_stft[MAX_DEV];

static int ensure_stft_tables(cudaStream_t st) {
    ZS_TRY(ensure_device());
    StftTables& t = g_stft[t_dev];
    if (t.w1024 && t.win) return ZS_OK;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (t.w1024 && t.win) return ZS_OK;
    std::vector<float2> w(1024);
    std::vector<float> win(ST_WIN);
    const double two_pi = 6.283185307179586476925286766559;
    for (int m = 0; m < 1024; ++m) w[m] = make_float2(static_cast<float>(cos(two_pi * m / 1024.0)), static_cast<float>(-sin(two_pi * m / 1024.0)));
    for (int m = 0; m < ST_WIN; ++m) win[m] = static_cast<float>(0.5 - 0.5 * cos(two_pi * m / ST_WIN));   // scipy get_window('hann', 800, fftbins=True)
    CUDA_TRY(cudaMalloc(&t.w1024, 1024 * sizeof(float2)));
    CUDA_TRY(cudaMalloc(&t.win, ST_WIN * sizeof(float)));
    CUDA_TRY(cudaMemcpy(t.w1024, w.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.win, win.data(), ST_WIN * sizeof(float), cudaMemcpyHostToDevice));
    (void)st;
    return ZS_OK;
}

extern "C" int zs_stft_tile_frames(void) { return GL_F; }

extern "C" size_t zs_griffin_lim_workspace_bytes(long long total_frames, long long total_samples) {
    if (total_frames < 1) return 0;
    const size_t spec = align256(static_cast<size_t>(total_frames) * ST_NBIN * sizeof(float2));
    return 2 * spec + align256(static_cast<size_t>(total_frames) * ST_NBIN * sizeof(float)) + align256(static_cast<size_t>(std::max(total_samples, 1LL)) * sizeof(float));
}

static StftMeta stft_meta(const int32_t* meta, int U) {
    StftMeta m;
    m.frame_start = meta; m.sample_start = meta + (U + 1); m.tile_start = meta + 2 * (U + 1);
    return m;
}

extern "C" int zs_griffin_lim(const float* spec, const int32_t* meta, int U, long long total_frames, long long total_samples, int total_tiles,
                              int n_iter, float preemphasis, float* wav, void* workspace, size_t workspace_bytes, void* stream) {
    if (!spec || !meta || !wav) return fail(ZS_ERR_ARG, "griffin_lim: null argument");
    if (U < 1 || total_frames < 4 || total_tiles < 1 || n_iter < 0) return fail(ZS_ERR_ARG, "griffin_lim: U %d, %lld frames, %d tiles, %d iterations", U, total_frames, total_tiles, n_iter);
    if (total_frames * ST_NBIN >= (1LL << 31)) return fail(ZS_ERR_ARG, "griffin_lim: %lld frames exceed the 32-bit element offsets of a call (split the batch)", total_frames);
    const size_t need = zs_griffin_lim_workspace_bytes(total_frames, total_samples);
    if (!workspace || workspace_bytes < need) return fail(ZS_ERR_WORKSPACE, "griffin_lim: workspace %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ZS_TRY(ensure_stft_tables(st));
    Carver c(workspace);
    const size_t n_el = static_cast<size_t>(total_frames) * ST_NBIN;
    float2* xa = static_cast<float2*>(c.take(n_el * sizeof(float2)));
    float2* xb = static_cast<float2*>(c.take(n_el * sizeof(float2)));
    float* mag = static_cast<float*>(c.take(n_el * sizeof(float)));
    float* raw = static_cast<float*>(c.take(static_cast<size_t>(std::max(total_samples, 1LL)) * sizeof(float)));
    {
        LaunchScope scope(st, KC_OTHER, 0.0, "denormalise_kernel");
        denormalise_kernel<<<static_cast<unsigned>((n_el + 255) / 256), 256, 0, st>>>(spec, mag, n_el, 100.f, 20.f);     // hps/hps.py:31-32
        CUDA_TRY(cudaGetLastError());
    }
    ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(gl_iter_kernel), GL_SMEM_BYTES));
    GlParams p;
    memset(&p, 0, sizeof(p));
    p.mag = mag; p.meta = stft_meta(meta, U); p.U = U; p.w1024 = g_stft[t_dev].w1024; p.win = g_stft[t_dev].win; p.wav = raw;
    // 5 N log2 N flops per complex FFT of 512 points; a tile runs GL_INV inverse and GL_F forward transforms
    const double fft_flops = 5.0 * 512 * 9 + 6.0 * 513;
    for (int it = 0; it <= n_iter; ++it) {      // convert.py:45-51: n_iter x (istft, stft, projection), then one more istft
        p.first = it == 0; p.final = it == n_iter;
        p.x_in = (it & 1) ? xa : xb;      // iteration `it` reads what iteration it-1 wrote (iteration 0 reads the magnitudes)
        p.x_out = (it & 1) ? xb : xa;
        LaunchScope scope(st, KC_OTHER, fft_flops * (p.final ? GL_INV : GL_INV + GL_F) * total_tiles, p.final ? "gl_iter_kernel final" : "gl_iter_kernel");
        gl_iter_kernel<<<total_tiles, GL_THREADS, GL_SMEM_BYTES, st>>>(p);
        CUDA_TRY(cudaGetLastError());
    }
    {   // convert.py:60 de-pre-emphasis
        LaunchScope scope(st, KC_OTHER, 0.0, "deemphasis_kernel");
        // grid.x covers the longest utterance: bounded by total_samples; threads past an utterance's end return at once
        const unsigned gx = static_cast<unsigned>((std::max(total_samples, 1LL) + 64 * 128 - 1) / (64 * 128));
        deemphasis_kernel<<<dim3(gx, U), 128, 0, st>>>(raw, wav, p.meta.sample_start, U, preemphasis);
        CUDA_TRY(cudaGetLastError());
    }
    return ZS_OK;
}

extern "C" int zs_frame_power(const float* wav, const int32_t* sample_start, const int32_t* pframe_start, int U, int max_frames,
                              float* power, void* stream) {
    if (!wav || !sample_start || !pframe_start || !power || U < 1 || max_frames < 1) return fail(ZS_ERR_ARG, "frame_power: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LaunchScope scope(st, KC_OTHER, 0.0, "frame_power_kernel");
    frame_power_kernel<<<dim3((max_frames + 7) / 8, U), 256, 0, st>>>(wav, sample_start, pframe_start, U, power);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}

extern "C" int zs_spectrogram(const float* wav, const int32_t* meta, int U, int total_tiles, float preemphasis, float* spec32,
                              void* spec16, void* stream) {
    if (!wav || !meta || (!spec32 && !spec16)) return fail(ZS_ERR_ARG, "spectrogram: null argument");
    if (U < 1 || total_tiles < 1) return fail(ZS_ERR_ARG, "spectrogram: U %d, %d tiles", U, total_tiles);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ZS_TRY(ensure_stft_tables(st));
    ZS_TRY(set_smem_attr(reinterpret_cast<const void*>(spec_kernel), GL_SMEM_BYTES));
    SpecParams p;
    memset(&p, 0, sizeof(p));
    p.wav = wav; p.spec32 = spec32; p.spec16 = static_cast<__half*>(spec16); p.meta = stft_meta(meta, U); p.U = U;
    p.w1024 = g_stft[t_dev].w1024; p.win = g_stft[t_dev].win; p.preemph = preemphasis; p.max_db = 100.f; p.ref_db = 20.f;
    LaunchScope scope(st, KC_OTHER, (5.0 * 512 * 9 + 6.0 * 513) * GL_F * total_tiles, "spec_kernel");
    spec_kernel<<<total_tiles, GL_THREADS, GL_SMEM_BYTES, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return ZS_OK;
}
