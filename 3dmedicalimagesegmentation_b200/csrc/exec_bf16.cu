#include "exec_impl.cuh"
namespace b200 { ExecIface* make_exec_bf16(const UnetrConfig& c) { return new ExecImpl<bf16>(c); } }
