// Shared between api_ops.cu and engine.cu.
#pragma once
namespace madm {
void set_global_error(const char* msg);
const char* global_error();
}  // namespace madm
