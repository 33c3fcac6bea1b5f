// hpfw_b200/csrc/cqt.cu — stage 1 (CQT front end). PLACEHOLDER while the kernels are being written: every entry point
// reports HPFW_ERR_STATE (it does NOT fall back to any CPU path).
#include "common.cuh"

extern "C" {
int hpfw_cqt_cols(int64_t) { return 0; }
int hpfw_cqt_spectrogram(hpfw_ctx *, const float *, int64_t, float *, int *) {
    HPFW_FAIL(HPFW_ERR_STATE, "CQT kernels not built into this library yet");
}
int hpfw_cqt_spectrogram_device(hpfw_ctx *, const float *, int64_t, float *, void *) {
    HPFW_FAIL(HPFW_ERR_STATE, "CQT kernels not built into this library yet");
}
int hpfw_cqt_magnitude(hpfw_ctx *, const float *, int64_t, float *, int *) {
    HPFW_FAIL(HPFW_ERR_STATE, "CQT kernels not built into this library yet");
}
int hpfw_hashprint_words_for_samples(int64_t) { return 0; }
int hpfw_calc_hashprint_audio(hpfw_ctx *, const float *, int64_t, uint64_t *, int *) {
    HPFW_FAIL(HPFW_ERR_STATE, "CQT kernels not built into this library yet");
}
int hpfw_calc_hashprint_audio_device(hpfw_ctx *, const float *, int64_t, uint64_t *, void *) {
    HPFW_FAIL(HPFW_ERR_STATE, "CQT kernels not built into this library yet");
}
}
