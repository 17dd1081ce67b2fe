#include "exec_impl.cuh"
namespace b200 { ExecIface* make_exec_f32(const UnetrConfig& c) { return new ExecImpl<float>(c); } }
