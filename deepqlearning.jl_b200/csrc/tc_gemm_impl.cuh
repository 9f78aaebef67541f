// tc_gemm_impl.cuh - tcgen05 path (placeholder until the tensor-core kernels land: every entry declines).
#pragma once
namespace {
bool tc_conv_fwd(dqn_engine*, const char*, const dqn::ConvFwdOp&, double, double) { return false; }
bool tc_dense_fwd(dqn_engine*, const char*, const dqn::DenseFwdOp*, int, double, double) { return false; }
bool tc_dense_wgrad(dqn_engine*, const char*, const dqn::DenseWgradOp*, int, double, double) { return false; }
bool tc_dense_dgrad(dqn_engine*, const char*, const dqn::DenseDgradOp&, double, double) { return false; }
bool tc_conv_wgrad(dqn_engine*, const char*, const dqn::ConvWgradOp&, double, double) { return false; }
bool tc_conv_dgrad(dqn_engine*, const char*, const dqn::ConvDgradOp&, double, double) { return false; }
void tc_init(dqn_engine*) {}
void tc_destroy(dqn_engine*) {}
void tc_params_changed(dqn_engine*) {}
}
