// Definitions of the opaque C-ABI handles (shared by the capi_*.cu translation units).
#pragma once
#include "../../include/ligero_b200.h"
#include "lg_internal.h"

struct lg_ctx {
  lg::Ctx c;
  bool col_len_prefix = true;
  bool leaf_len_prefix = true;
};
struct lg_matrix {
  lg::Matrix m;
  lg_ctx* owner = nullptr;
};
struct lg_constraints {
  lg_ctx* owner = nullptr;
  size_t mk = 0, nnz = 0, n_consts = 0;
  uint32_t* col_ptr = nullptr;   // mk + 1
  uint32_t* row_idx = nullptr;   // nnz, row of A in [0, 4mk)
  uint32_t* val_id = nullptr;    // nnz: 0 -> +1, 1 -> -1, v >= 2 -> consts[v - 2]
  lg::Fr* consts = nullptr;
};

namespace lg {
bool is_device_ptr(const void* p);
}

// constraints.cu: CSC of the right-hand block of A built on the device from host node arrays
namespace lg {
int build_constraints_device(lg_ctx* ctx, const uint8_t* type, const uint32_t* l, const uint32_t* r, size_t n, const uint32_t* vidp,
                             const uint32_t* vidn, size_t n_const_values, const uint32_t* outputs, size_t n_out, size_t mk,
                             const uint64_t* const_table, size_t n_table, lg_constraints** out);
}
