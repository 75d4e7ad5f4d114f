// rlca_conv_tc.cuh — internal interface of the tcgen05 conv tower (rlca_conv_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

int rlca_conv_tc_init();
size_t rlca_conv_tc_image_floats();
// builds the pre-swizzled tf32 hi/lo operand image of both towers' conv weights (index 0 actor, 1 critic)
void rlca_conv_tc_prep(const float *const cv1w[2], const float *const cv1b[2], const float *const cv2w[2],
                       const float *const cv2b[2], float *img, cudaStream_t s);
// F = [2][nb][4096] relu(conv2(relu(conv1(obs)))) in flatten order c*128+q; Fs (optional) = [tower][hi,lo][nb][4096]
int rlca_conv_tc_forward(const float *obs, const float *img, float *F, float *Fs, int nb, int num_sms, cudaStream_t s);
