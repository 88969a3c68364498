// Driver of scripts/etl_fuzz.py: the ETL entry points on one (mutated) file, built with ASan + UBSan.
#include <cstdio>
#include "../omnidirectional_collaborative_filtering_b200/csrc/ocf_etl.cpp"
namespace ocf { std::string& last_error(){ static std::string e; return e;} int fail(int c, const std::string& m){ last_error()=m; return c;} }
// drv json <vocab> <file> <paired> | drv csv <file> <ncols> <outdir>
int main(int argc, char** argv) {
  std::string mode = argv[1];
  if (mode == "json") {
    ocf_vocab* v = nullptr;
    if (ocf_vocab_load_json(argv[2], &v)) { printf("vocab: %s\n", ocf::last_error().c_str()); return 0; }
    ocf_ratings* r = nullptr;
    int paired = atoi(argv[4]);
    if (ocf_ratings_load_json(argv[3], v, paired, &r)) { printf("err: %s\n", ocf::last_error().c_str()); ocf_vocab_destroy(v); return 0; }
    int64_t info[4]; ocf_ratings_info(r, info);
    std::vector<char> kb(info[1] + 1); std::vector<int64_t> off(info[0] + 1);
    ocf_ratings_keys(r, kb.data(), off.data());
    for (int w = 0; w <= paired; ++w) {
      std::vector<int64_t> rp(info[0] + 1); std::vector<int32_t> c(info[2 + w] + 1); std::vector<float> x(info[2 + w] + 1); std::vector<uint8_t> none(info[0] + 1);
      ocf_ratings_csr(r, w, rp.data(), c.data(), x.data(), none.data());
      if (rp[info[0]] != info[2 + w]) { printf("BUG nnz mismatch\n"); return 1; }
    }
    printf("ok rows %lld\n", (long long)info[0]);
    ocf_ratings_destroy(r); ocf_vocab_destroy(v);
  } else {
    ocf_csv* c = nullptr;
    if (ocf_csv_load(argv[2], atoi(argv[3]), &c)) { printf("err: %s\n", ocf::last_error().c_str()); return 0; }
    int64_t n; ocf_csv_rows(c, &n);
    std::vector<int64_t> order(n); for (int64_t k = 0; k < n; ++k) order[k] = (k * 7 + 3) % (n ? n : 1);
    if (n > 0 && n % 7 == 0) for (int64_t k = 0; k < n; ++k) order[k] = n - 1 - k;
    double fr[3] = {.8, .1, .1};
    for (int ts = 0; ts < 2; ++ts) for (int rev = 0; rev < 2; ++rev) for (int cast = 0; cast < 2; ++cast) {
      int rc = ocf_split_write(c, order.data(), n, fr, argv[4], cast, 1, ts, 1, rev);
      if (rc) printf("split err: %s\n", ocf::last_error().c_str());
    }
    for (int rev = 0; rev < 2; ++rev) for (int cast = 0; cast < 2; ++cast) {
      ocf_split* sp = nullptr;
      if (ocf_split_build(c, order.data(), n, fr, cast, rev, &sp)) { printf("build err: %s\n", ocf::last_error().c_str()); continue; }
      int64_t info[13]; ocf_split_info(sp, info);
      for (int set = 0; set < 3; ++set) {
        std::vector<char> kb(info[2 + 4 * set] + 1); std::vector<int64_t> off(info[1 + 4 * set] + 1);
        ocf_split_keys(sp, set, kb.data(), off.data());
        for (int part = 0; part < (set ? 2 : 1); ++part) {
          int64_t nnz = info[3 + 4 * set + part];
          std::vector<int64_t> rp(info[1 + 4 * set] + 1); std::vector<int32_t> cc(nnz + 1); std::vector<float> vv(nnz + 1); std::vector<uint8_t> none(info[1 + 4 * set] + 1);
          ocf_split_csr(sp, set, part, rp.data(), cc.data(), vv.data(), none.data());
          if (rp[info[1 + 4 * set]] != nnz) { printf("BUG split nnz mismatch\n"); return 1; }
          for (int64_t k = 0; k < nnz; ++k) if (cc[k] < 0 || cc[k] >= info[0]) { printf("BUG column out of range\n"); return 1; }
        }
      }
      int64_t need = 0; ocf_split_columns_json(sp, nullptr, 0, &need);
      std::vector<char> txt(need + 1); ocf_split_columns_json(sp, txt.data(), need, &need);
      ocf_split_destroy(sp);
    }
    printf("ok rows %lld\n", (long long)n);
    ocf_csv_destroy(c);
  }
  return 0;
}
