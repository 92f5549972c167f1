// C++ driver for the host layer include/vdf_host.hpp: replays fixtures written by tests/test_cpp_host.py
// (inputs + oracle answers) through the reference-shaped C++ interface and compares bytes.
//   usage: host_check <fixture-dir>      exit 0 = all equal, 2 = mismatch, 3 = library/GPU error
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>

#include "../../include/vdf_host.hpp"

using namespace vdf_host;

static std::vector<uint8_t> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
template <class T>
static std::vector<T> as(const std::vector<uint8_t>& raw) {
  std::vector<T> v(raw.size() / sizeof(T));
  std::memcpy(v.data(), raw.data(), v.size() * sizeof(T));
  return v;
}

int main(int argc, char** argv) {
  if (argc != 2) { std::fprintf(stderr, "usage: host_check <fixture-dir>\n"); return 1; }
  const std::string d = std::string(argv[1]) + "/";
  int bad = 0;
  try {
    init(0);
    {  // MinRootVDF::check over ragged chains (src/minroot.rs:369-371), then Evaluation::append (:428-438)
      auto results = as<State>(slurp(d + "mr_results.bin")), originals = as<State>(slurp(d + "mr_originals.bin"));
      auto ts = as<uint64_t>(slurp(d + "mr_t.bin"));
      auto want = slurp(d + "mr_ok.bin");
      auto got = PallasVDF::check_batch(results, ts, originals);
      for (size_t k = 0; k < got.size(); k++) bad += (got[k] != (want[k] != 0));
      auto inv = PallasVDF::inverse_eval_batch(results, 6);
      auto want_inv = as<State>(slurp(d + "mr_inverse6.bin"));
      for (size_t k = 0; k < inv.size(); k++) bad += !(inv[k] == want_inv[k]);
      auto seg = as<State>(slurp(d + "mr_segments.bin"));  // s0, s1 = eval(s0, 4), s2 = eval(s1, 4)
      Evaluation<PallasVDF> first{seg[1], 4}, second{seg[2], 4};
      auto joined = first.append(second);
      bad += !(joined && joined->t == 8 && joined->verify(seg[0]));
      Evaluation<PallasVDF> wrong{seg[0], 4};
      bad += first.append(wrong).has_value();
      std::printf("minroot: %zu chains, mismatches so far %d\n", got.size(), bad);
    }
    {  // commit() with resident generators and pasta_msm with travelling points
      auto k0d = as<Fe>(slurp(d + "msm_k0_d.bin"));
      auto scalars = as<Fe>(slurp(d + "msm_scalars.bin"));
      auto want = as<Point>(slurp(d + "msm_point.bin"));
      Generators g(VDFGPU_PALLAS, k0d[0], k0d[1], scalars.size(), true);
      bad += !(g.commit(scalars) == want[0]);
      auto pts = as<Affine>(slurp(d + "msm_points.bin"));
      bad += !(pasta_msm(VDFGPU_PALLAS, pts, scalars) == want[0]);
      std::printf("msm: n = %zu, mismatches so far %d\n", scalars.size(), bad);
    }
    {  // R1CSShape::multiply_vec, commit_T's T, fold, and one device-resident NIFS step
      auto dims = as<uint64_t>(slurp(d + "r1cs_dims.bin"));  // cons, vars, io, nnzA, nnzB, nnzC
      CooMatrix M[3];
      const char* names[3] = {"a", "b", "c"};
      for (int m = 0; m < 3; m++) {
        M[m].rows = as<uint64_t>(slurp(d + "r1cs_" + names[m] + "_rows.bin"));
        M[m].cols = as<uint64_t>(slurp(d + "r1cs_" + names[m] + "_cols.bin"));
        M[m].vals = as<Fe>(slurp(d + "r1cs_" + names[m] + "_vals.bin"));
      }
      R1CSShape shape(VDFGPU_FQ, dims[0], dims[1], dims[2], M[0], M[1], M[2]);
      auto W1 = as<Fe>(slurp(d + "r1cs_W1.bin")), W2 = as<Fe>(slurp(d + "r1cs_W2.bin"));
      auto X1 = as<Fe>(slurp(d + "r1cs_X1.bin")), X2 = as<Fe>(slurp(d + "r1cs_X2.bin"));
      auto u1 = as<Fe>(slurp(d + "r1cs_u1.bin")), r = as<Fe>(slurp(d + "r1cs_r.bin"));
      std::vector<Fe> z = W1;
      z.push_back(u1[0]);
      z.insert(z.end(), X1.begin(), X1.end());
      auto p = shape.multiply_vec(z);
      auto wantABC = as<Fe>(slurp(d + "r1cs_ABC.bin"));
      for (size_t k = 0; k < dims[0]; k++)
        bad += !(p.Az[k] == wantABC[k] && p.Bz[k] == wantABC[dims[0] + k] && p.Cz[k] == wantABC[2 * dims[0] + k]);
      auto k0d = as<Fe>(slurp(d + "msm_k0_d.bin"));
      Generators g(VDFGPU_PALLAS, k0d[0], k0d[1], dims[0] > dims[1] ? dims[0] : dims[1], true);
      auto [T, commT] = shape.commit_T(g, W1, u1[0], X1, W2, X2);
      auto wantT = as<Fe>(slurp(d + "r1cs_T.bin"));
      for (size_t k = 0; k < dims[0]; k++) bad += !(T[k] == wantT[k]);
      bad += !(commT == as<Point>(slurp(d + "r1cs_commT.bin"))[0]);
      std::vector<Fe> E1(dims[0], Fe{});
      RunningWitness run(shape, g);
      run.set(W1, E1, u1[0], X1);
      auto c = run.commit(W2, X2);
      bad += !(c.comm_T == commT);
      run.fold(r[0]);
      std::vector<Fe> Wf, Ef, Xf;   // get() sizes its outputs (empty vectors used to overflow the heap)
      Fe uf{};
      run.get(Wf, Ef, uf, Xf);
      bad += !(Wf.size() == dims[1] && Ef.size() == dims[0] && Xf.size() == dims[2]);
      // nova-snark answers a wrong length with InvalidWitnessLength; the host layer must not let the C ABI read past it
      auto throws = [](auto&& fn) { try { fn(); } catch (const std::invalid_argument&) { return 1; } return 0; };
      std::vector<Fe> shortW(W2.begin(), W2.end() - 1);
      bad += !throws([&] { run.commit(shortW, X2); });
      bad += !throws([&] { run.set(shortW, E1, u1[0], X1); });
      bad += !throws([&] { shape.commit_T(g, shortW, u1[0], X1, W2, X2); });
      bad += !throws([&] { shape.multiply_vec(shortW); });
      auto wantW = as<Fe>(slurp(d + "r1cs_Wfold.bin")), wantE = as<Fe>(slurp(d + "r1cs_Efold.bin"));
      for (size_t k = 0; k < dims[1]; k++) bad += !(Wf[k] == wantW[k]);
      for (size_t k = 0; k < dims[0]; k++) bad += !(Ef[k] == wantE[k]);
      fold(VDFGPU_FQ, W1, W2, E1, T, r[0]);   // host-vector variant must agree with the resident one
      for (size_t k = 0; k < dims[1]; k++) bad += !(W1[k] == wantW[k]);
      for (size_t k = 0; k < dims[0]; k++) bad += !(E1[k] == wantE[k]);
      std::printf("r1cs: cons = %llu, mismatches so far %d\n", (unsigned long long)dims[0], bad);
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "host_check: %s\n", e.what());
    return 3;
  }
  if (bad) { std::printf("MISMATCH (%d)\n", bad); return 2; }
  std::printf("OK\n");
  return 0;
}
