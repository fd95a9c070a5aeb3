// Shim: the reference includes <XoshiroCpp.hpp> (Reputeless/Xoshiro-cpp v1.1,
// fetched by CPM at configure time, not vendored, no network here). This maps
// the one class the reference uses onto our own xoshiro256++ engine.
#pragma once
#include "../../qkd_ldpc_v_b200/host/xoshiro256pp.hpp"
namespace XoshiroCpp {
using Xoshiro256PlusPlus = ::qkdldpc::Xoshiro256pp;
}
