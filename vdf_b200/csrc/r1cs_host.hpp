// COO -> CSR conversion of an R1CS shape on the host (pure C++; shared by the C ABI and tests/emul).
// nova-snark stores A, B, C as COO triples (row: usize, col: usize, val: Scalar) [SURVEY.md 8a row a5];
// the device wants one CSR with 3*cons rows: A rows, then B rows, then C rows.  Stable within a row, so
// duplicate (row, col) entries keep their order (their products are summed by the kernels anyway).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace vdf {

struct HostCsr {
  std::vector<uint32_t> row_ptr, col;
  std::vector<uint8_t> val;  // 32 bytes per entry
};

inline HostCsr coo_to_csr(size_t num_cons, size_t ncols, const uint64_t* const rows[3], const uint64_t* const cols[3],
                          const uint8_t* const vals[3], const size_t nnzs[3]) {
  size_t nnz = nnzs[0] + nnzs[1] + nnzs[2];
  for (int m = 0; m < 3; m++) {
    if (nnzs[m] && (!rows[m] || !cols[m] || !vals[m])) throw std::invalid_argument("r1cs: null COO array");
    for (size_t k = 0; k < nnzs[m]; k++)
      if (rows[m][k] >= num_cons || cols[m][k] >= ncols) throw std::invalid_argument("r1cs: COO index out of range");
  }
  const size_t R = 3 * num_cons;
  HostCsr out;
  out.row_ptr.assign(R + 1, 0);
  out.col.resize(nnz);
  out.val.resize(nnz * 32);
  for (int m = 0; m < 3; m++)
    for (size_t k = 0; k < nnzs[m]; k++) out.row_ptr[m * num_cons + rows[m][k] + 1]++;
  for (size_t r = 0; r < R; r++) out.row_ptr[r + 1] += out.row_ptr[r];
  std::vector<uint32_t> cursor(out.row_ptr.begin(), out.row_ptr.end() - 1);
  for (int m = 0; m < 3; m++)
    for (size_t k = 0; k < nnzs[m]; k++) {
      uint32_t pos = cursor[m * num_cons + rows[m][k]]++;
      out.col[pos] = (uint32_t)cols[m][k];
      std::memcpy(&out.val[(size_t)pos * 32], vals[m] + k * 32, 32);
    }
  return out;
}

// Column view of the same matrices for the transposed product of Spartan's inner sum-check (nova-snark
// compute_eval_table_sparse [R]): columns in order, the entries of a column in CSR order, each with its stacked row
// (matrix * cons + row) and its own copy of the value (sequential reads).  Columns with more than HEAVY entries -- the
// constant column of the step circuit holds one per round -- are listed: a warp sums each of them.
struct HostCsc {
  static constexpr uint32_t HEAVY = 64;
  std::vector<uint32_t> col_ptr, srow, heavy;
  std::vector<uint8_t> val;  // 32 bytes per entry, column order
};

inline HostCsc csr_to_csc(const HostCsr& csr, size_t ncols) {
  const size_t nnz = csr.col.size(), R = csr.row_ptr.size() - 1;
  HostCsc out;
  out.col_ptr.assign(ncols + 1, 0);
  out.srow.resize(nnz);
  out.val.resize(nnz * 32);
  for (size_t k = 0; k < nnz; k++) out.col_ptr[csr.col[k] + 1]++;
  for (size_t c = 0; c < ncols; c++) {
    if (out.col_ptr[c + 1] > HostCsc::HEAVY) out.heavy.push_back((uint32_t)c);
    out.col_ptr[c + 1] += out.col_ptr[c];
  }
  std::vector<uint32_t> cursor(out.col_ptr.begin(), out.col_ptr.end() - 1);
  for (size_t r = 0; r < R; r++)
    for (uint32_t k = csr.row_ptr[r]; k < csr.row_ptr[r + 1]; k++) {
      const uint32_t pos = cursor[csr.col[k]]++;
      out.srow[pos] = (uint32_t)r;
      std::memcpy(&out.val[(size_t)pos * 32], &csr.val[(size_t)k * 32], 32);
    }
  return out;
}

}  // namespace vdf
