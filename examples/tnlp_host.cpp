// A compiled NLP host over the C ABI, no Python in the process: the method set of IPOPT's
// Ipopt::TNLP (get_nlp_info / eval_f / eval_grad_f / eval_g / eval_jac_g / eval_h, the C++
// interface behind the cyipopt object of pycollo/nlp.py:36-76) implemented on include/pcx.h.
// IPOPT itself is not in the image, so the class does not derive from Ipopt::TNLP; the
// signatures are TNLP's (Index = int, Number = double), and main() plays the solver's part:
// one call of every callback at the iterate in <x.bin>, results written to <out.bin>
// (tests/test_cabi_and_host.py compares them with the Python engine, bit for bit).
//
//   g++ -std=c++17 -Iinclude examples/tnlp_host.cpp -Lpycollo_b200 -lpcx -Wl,-rpath,$PWD/pycollo_b200 -o tnlp_host
//   ./tnlp_host problem.pcxspec x.bin lam.bin out.bin        (spec: Engine.save_spec / Engine.write_spec)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "pcx.h"

typedef int Index;
typedef double Number;

class PcxTNLP {
public:
    PcxTNLP(const char* spec_path, int device) {
        if (pcx_create_from_file(spec_path, device, &eng_) != PCX_OK) {
            fprintf(stderr, "pcx_create_from_file: %s\n", pcx_last_error(nullptr));
            eng_ = nullptr;
            return;
        }
        pcx_sizes(eng_, &n_, &m_, nullptr, &nnz_j_, &nnz_h_, nullptr);
        // IPOPT wants triplets in any fixed order: the row-major Jacobian / lower-triangular
        // Hessian of the reference's cyipopt plumbing (iteration.py:930-933, 965-968); the
        // engine's native order is CasADi's CCS, perm maps one onto the other
        jr_.resize(nnz_j_); jc_.resize(nnz_j_); jp_.resize(nnz_j_);
        hr_.resize(nnz_h_); hc_.resize(nnz_h_); hp_.resize(nnz_h_);
        ok_ = pcx_structure_jac(eng_, PCX_ORDER_ROW_MAJOR, jr_.data(), jc_.data(), jp_.data()) == PCX_OK
           && pcx_structure_hess(eng_, PCX_ORDER_ROW_MAJOR, hr_.data(), hc_.data(), hp_.data()) == PCX_OK;
        jv_.resize(nnz_j_); hv_.resize(nnz_h_);
    }
    ~PcxTNLP() { if (eng_) pcx_destroy(eng_); }
    bool ok() const { return eng_ && ok_; }

    bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag) const {
        n = (Index)n_; m = (Index)m_; nnz_jac_g = (Index)nnz_j_; nnz_h_lag = (Index)nnz_h_;
        return true;
    }
    // f, grad f and g of one iterate leave in ONE launch; IPOPT asks for them one by one
    // with new_x = false after the first, so they are cached per iterate
    bool eval_f(Index, const Number* x, bool new_x, Number& obj) { return head(x, new_x) && ((obj = f_), true); }
    bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad) {
        if (!head(x, new_x)) return false;
        memcpy(grad, grad_.data(), sizeof(Number) * (size_t)n);
        return true;
    }
    bool eval_g(Index, const Number* x, bool new_x, Index m, Number* g) {
        if (!head(x, new_x)) return false;
        memcpy(g, c_.data(), sizeof(Number) * (size_t)m);
        return true;
    }
    bool eval_jac_g(Index, const Number* x, bool, Index, Index nele, Index* iRow, Index* jCol, Number* values) {
        if (!values) {                                   // structure call
            for (Index k = 0; k < nele; ++k) { iRow[k] = (Index)jr_[k]; jCol[k] = (Index)jc_[k]; }
            return true;
        }
        if (pcx_eval_jac(eng_, x, jv_.data(), PCX_HOST, nullptr) != PCX_OK) return false;
        for (Index k = 0; k < nele; ++k) values[k] = jv_[(size_t)jp_[k]];
        return true;
    }
    bool eval_h(Index, const Number* x, bool, Number obj_factor, Index, const Number* lambda, bool,
                Index nele, Index* iRow, Index* jCol, Number* values) {
        if (!values) {
            for (Index k = 0; k < nele; ++k) { iRow[k] = (Index)hr_[k]; jCol[k] = (Index)hc_[k]; }
            return true;
        }
        if (pcx_eval_hess(eng_, x, lambda, &obj_factor, hv_.data(), PCX_HOST, nullptr) != PCX_OK) return false;
        for (Index k = 0; k < nele; ++k) values[k] = hv_[(size_t)hp_[k]];
        return true;
    }

private:
    bool head(const Number* x, bool new_x) {
        if (!new_x && have_) return true;
        grad_.resize(n_); c_.resize(m_);
        have_ = pcx_eval(eng_, PCX_EVAL_F | PCX_EVAL_GRAD | PCX_EVAL_C, x, nullptr, nullptr, &f_,
                         grad_.data(), c_.data(), nullptr, nullptr, nullptr, PCX_HOST, nullptr) == PCX_OK;
        if (!have_) fprintf(stderr, "pcx_eval: %s\n", pcx_last_error(eng_));
        return have_;
    }
    pcx_engine* eng_ = nullptr;
    bool ok_ = false, have_ = false;
    int64_t n_ = 0, m_ = 0, nnz_j_ = 0, nnz_h_ = 0;
    std::vector<int64_t> jr_, jc_, jp_, hr_, hc_, hp_;
    std::vector<Number> jv_, hv_, grad_, c_;
    Number f_ = 0.0;
};

static bool read_doubles(const char* path, std::vector<double>& v, size_t n) {
    v.resize(n);
    FILE* fh = fopen(path, "rb");
    if (!fh) return false;
    const size_t got = fread(v.data(), sizeof(double), n, fh);
    fclose(fh);
    return got == n;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s spec x.bin lam.bin out.bin\n", argv[0]); return 2; }
    PcxTNLP nlp(argv[1], 0);
    if (!nlp.ok()) return 1;
    Index n, m, nj, nh;
    nlp.get_nlp_info(n, m, nj, nh);
    std::vector<double> x, lam;
    if (!read_doubles(argv[2], x, (size_t)n) || !read_doubles(argv[3], lam, (size_t)m)) {
        fprintf(stderr, "bad x / lam file\n");
        return 1;
    }
    std::vector<Index> jr(nj), jc(nj), hr(nh), hc(nh);
    std::vector<double> grad(n), g(m), jv(nj), hv(nh);
    double f = 0.0;
    bool ok = nlp.eval_jac_g(n, nullptr, false, m, nj, jr.data(), jc.data(), nullptr)
           && nlp.eval_h(n, nullptr, false, 1.0, m, nullptr, false, nh, hr.data(), hc.data(), nullptr)
           && nlp.eval_f(n, x.data(), true, f)
           && nlp.eval_grad_f(n, x.data(), false, grad.data())
           && nlp.eval_g(n, x.data(), false, m, g.data())
           && nlp.eval_jac_g(n, x.data(), false, m, nj, nullptr, nullptr, jv.data())
           && nlp.eval_h(n, x.data(), false, 0.75, m, lam.data(), true, nh, nullptr, nullptr, hv.data());
    if (!ok) return 1;
    FILE* out = fopen(argv[4], "wb");
    if (!out) return 1;
    const int64_t dims[4] = {n, m, nj, nh};
    fwrite(dims, sizeof(int64_t), 4, out);
    fwrite(&f, sizeof(double), 1, out);
    fwrite(grad.data(), sizeof(double), grad.size(), out);
    fwrite(g.data(), sizeof(double), g.size(), out);
    fwrite(jr.data(), sizeof(Index), jr.size(), out);
    fwrite(jc.data(), sizeof(Index), jc.size(), out);
    fwrite(jv.data(), sizeof(double), jv.size(), out);
    fwrite(hr.data(), sizeof(Index), hr.size(), out);
    fwrite(hc.data(), sizeof(Index), hc.size(), out);
    fwrite(hv.data(), sizeof(double), hv.size(), out);
    fclose(out);
    printf("n=%d m=%d nnz_jac=%d nnz_hess=%d f=%.17g\n", n, m, nj, nh, f);
    return 0;
}
