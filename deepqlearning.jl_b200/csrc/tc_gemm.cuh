// tc_gemm.cuh - entry points of the tcgen05 (DQN_MATH_3XTF32) contraction path.  Each returns false when it
// does not take the launch (math_mode fp32, or a shape the tensor-core kernels do not cover), in which case
// the caller launches the fp32 CUDA-core implicit GEMM of igemm.cuh.  Both are device paths.
#pragma once
#include "igemm.cuh"
struct dqn_engine;
namespace {
bool tc_conv1_eligible(dqn_engine* e, const dqn::ConvGeom& g);
void tc_conv1_init(dqn_engine* e);
bool tc_conv1_fwd(dqn_engine* e, const char* name, const dqn::ConvFwdOp& op, double flops, double bytes);
bool tc_conv_fwd(dqn_engine* e, const char* name, const dqn::ConvFwdOp* ops, int nops, double flops, double bytes);
bool tc_dense_fwd(dqn_engine* e, const char* name, const dqn::DenseFwdOp* ops, int nops, double flops, double bytes);
bool tc_dense_wgrad(dqn_engine* e, const char* name, const dqn::DenseWgradOp* ops, int ntow, double flops, double bytes);
bool tc_dense_dgrad(dqn_engine* e, const char* name, const dqn::DenseDgradOp& op, double flops, double bytes);
bool tc_dense_dgrad2(dqn_engine* e, const char* name, const dqn::DenseDgradOp* ops, int ntow, double flops, double bytes);
bool tc_conv_wgrad(dqn_engine* e, const char* name, const dqn::ConvWgradOp& op, double flops, double bytes);
bool tc_conv_dgrad_merged(dqn_engine* e, const char* name, const dqn::ConvDgradMergedOp& op, double flops, double bytes);
bool tc_conv_dgrad(dqn_engine* e, const char* name, const dqn::ConvDgradOp& op, double flops, double bytes);
void tc_init(dqn_engine* e);
void tc_destroy(dqn_engine* e);
void tc_params_changed(dqn_engine* e);
}
