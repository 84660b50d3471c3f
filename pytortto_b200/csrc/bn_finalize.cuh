// Per-channel BatchNorm finalisation math, shared by the single-GPU finalize kernels (batchnorm.cu) and by the fused
// "all-reduce the statistics over NVLink peer memory + finalise" kernels of the data-parallel path (comm.cu).
// Reference: BatchNorm.forward / backward, /root/reference/src/tortto/autograd/grad_nn.py:909-989.
#pragma once

namespace ttb {

struct BnFwdFinalize {  // inputs: s0 = sum(x), s1 = sum(x*x) over `count` elements of one channel
  double count;
  float eps, momentum, one_minus_momentum, unbias;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float *mean, *var_eps, *sd, *scale, *shift;

  __device__ __forceinline__ void operator()(int i, double s0, double s1) const {
    double mu = s0 / count;
    double var = s1 / count - mu * mu;  // biased variance, grad_nn.py:924
    if (var < 0.0) var = 0.0;
    float muf = (float)mu, varf = (float)var;
    if (running_mean) running_mean[i] = __fadd_rn(__fmul_rn(one_minus_momentum, running_mean[i]), __fmul_rn(momentum, muf));
    if (running_var)
      running_var[i] = __fadd_rn(__fmul_rn(one_minus_momentum, running_var[i]), __fmul_rn(momentum, __fmul_rn(varf, unbias)));
    float ve = __fadd_rn(varf, eps);
    float s = sqrtf(ve);
    mean[i] = muf;
    var_eps[i] = ve;
    sd[i] = s;
    float g = gamma ? gamma[i] : 1.f;
    scale[i] = g / s;
    shift[i] = beta ? beta[i] : 0.f;  // the additive term of y = (x - mean)*scale + beta (the kernels centre x first)
  }
};

struct BnBwdFinalize {  // inputs: sdy = sum(dy), sdyx = sum(dy*(x-mean)); c = channel count (coef is [3][c])
  double count;
  int c;
  const float* gamma;
  const float* var_eps;
  const float* sd;
  float* dgamma;
  float* dbeta;
  float* coef;

  __device__ __forceinline__ void operator()(int i, double sdy, double sdyx) const {
    if (dbeta) dbeta[i] = (float)sdy;
    if (dgamma) dgamma[i] = (float)(sdyx / (double)sd[i]);
    float g = gamma ? gamma[i] : 1.f;
    coef[i] = g / sd[i];                                             // c1
    coef[c + i] = (float)(sdy / count);                              // c2
    coef[2 * c + i] = (float)(sdyx / (count * (double)var_eps[i]));  // c3
  }
};

// host helper: N/(N-1) and (1-momentum) are evaluated in double like the reference's Python floats (grad_nn.py:927-930)
inline void bn_fwd_host_factors(int64_t count, float momentum, float* unbias, float* one_minus_momentum) {
  *unbias = count > 1 ? (float)((double)count / (double)(count - 1)) : 1.f;
  *one_minus_momentum = (float)(1.0 - (double)momentum);
}

}  // namespace ttb
