// zksc.hpp -- C++17 host-side mirror of the reference's Rust API for the sumcheck path, over the C ABI of zksc.h.
//
// The reference (aagbotemi/zk-cryptography) is a Rust workspace and there is no Rust toolchain in this image, so the
// host side above the C ABI is written in C++ (compiled code, like the reference) with the reference's names, argument
// meaning and error behaviour, so that tests/cpp/reference_cases.cpp reads like the reference's own #[test]s.  The same
// mapping as Rust source (uncompiled) is in rust/zksc-sumcheck.  Header only; link with libzksc.so.
//
//   reference item                                                       here
//   polynomial/src/multilinear/evaluation_form.rs  Multilinear<F>        zk::Multilinear
//   polynomial/src/composed/composed_multilinear.rs ComposedMultilinear  zk::ComposedMultilinear
//   polynomial/src/univariate/sparse_univariate.rs                       zk::SparseUnivariatePolynomial, zk::UnivariateMonomial
//   transcripts/fiat-shamir/src/fiat_shamir.rs                           zk::FiatShamirTranscript
//   sumcheck/src/sumcheck.rs                                             zk::Sumcheck, zk::SumcheckProof
//   sumcheck/src/composed/composed_sumcheck.rs                           zk::ComposedSumcheck, zk::composed::ComposedSumcheckProof
//   sumcheck/src/composed/multi_composed_sumcheck.rs                     zk::MultiComposedSumcheckProver / Verifier,
//                                                                        zk::ComposedSumcheckProof, zk::SubClaim
//   circuit/src/{gate,circuit}.rs, gkr/src/protocol.rs                   zk::Gate, zk::CircuitLayer, zk::Circuit, zk::GKRProtocol, zk::GKRProof
//
// Rust `new` is `new_` (keyword).  `assert!/panic!` sites throw zk::Error (code ZKSC_ERR_SHAPE); `Result<_, &'static str>`
// is zk::Result<T> with unwrap()/is_ok()/err.  Field elements are ark-ff's in-memory form (4 x u64 LE, Montgomery), so a
// std::vector<zk::Fr> is passed to the library without conversion.  All table-sized work runs on the GPU; there is no
// CPU fallback (without a CUDA device the first call throws).
#ifndef ZKSC_HPP
#define ZKSC_HPP
#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zksc.h"

namespace zk {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

template <class T>
struct Result {  // Result<T, &'static str>
    std::optional<T> ok;
    const char* err = nullptr;
    bool is_ok() const { return ok.has_value(); }
    const T& unwrap() const {
        if (!ok) throw Error(ZKSC_ERR_VERIFY, err ? err : "unwrap on Err");
        return *ok;
    }
};

// ark_test_curves::bls12_381::Fr -- Montgomery limbs, same bytes as ark-ff's Fp256
struct Fr {
    uint64_t v[4] = {0, 0, 0, 0};
    static Fr from(uint64_t x) { Fr r; zksc_fr_from_u64(x, r.v); return r; }
    static Fr zero() { return Fr(); }
    static Fr one() { return from(1); }
    static Fr from_be_bytes_mod_order(const uint8_t b[32]) { Fr r; zksc_fr_from_be_bytes_mod_order(b, r.v); return r; }
    Fr operator+(const Fr& o) const { Fr r; zksc_fr_add(v, o.v, r.v); return r; }
    Fr operator-(const Fr& o) const { Fr r; zksc_fr_sub(v, o.v, r.v); return r; }
    Fr operator*(const Fr& o) const { Fr r; zksc_fr_mul(v, o.v, r.v); return r; }
    Fr& operator+=(const Fr& o) { return *this = *this + o; }
    bool operator==(const Fr& o) const { return std::memcmp(v, o.v, 32) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    std::array<uint8_t, 32> to_bytes_be() const { std::array<uint8_t, 32> b; zksc_fr_to_be_bytes(v, b.data()); return b; }  // into_bigint().to_bytes_be()
};
static_assert(sizeof(Fr) == 32, "Fr must be 4 x u64");
inline const uint64_t* raw(const std::vector<Fr>& x) { return reinterpret_cast<const uint64_t*>(x.data()); }
inline uint64_t* raw(std::vector<Fr>& x) { return reinterpret_cast<uint64_t*>(x.data()); }
inline void append(std::vector<uint8_t>& out, const Fr& x) { auto b = x.to_bytes_be(); out.insert(out.end(), b.begin(), b.end()); }

// one process-wide context: on device 0, or the device ZKSC_DEVICE names, or -- ZKSC_DEVICES="0,1,2,3" -- ONE context over several devices
// of this process (zksc_ctx_create_multi: tables sharded over them, one host thread, one transcript; the reference's caller is one process)
class Context {
   public:
    static zksc_ctx* get() {
        static Context c;
        return c.h_;
    }
    static void check(int rc) {
        if (rc == ZKSC_OK) return;
        const char* m = zksc_last_error(get());
        throw Error(rc, std::string("zksc error ") + std::to_string(rc) + ": " + (m ? m : ""));
    }

   private:
    Context() {
        int rc;
        if (const char* many = std::getenv("ZKSC_DEVICES")) {
            std::vector<int> devs;
            for (const char* p = many; *p;) {
                devs.push_back(std::atoi(p));
                while (*p && *p != ',') p++;
                if (*p == ',') p++;
            }
            rc = zksc_ctx_create_multi(devs.data(), (int)devs.size(), &h_);
        } else {
            const char* d = std::getenv("ZKSC_DEVICE");
            rc = zksc_ctx_create(d ? std::atoi(d) : 0, &h_);
        }
        if (rc != ZKSC_OK) throw Error(rc, std::string("zksc_ctx_create: ") + zksc_last_error(nullptr) + " (there is no CPU fallback)");
    }
    ~Context() { zksc_ctx_destroy(h_); }
    zksc_ctx* h_ = nullptr;
};

// device tables of one proof, freed on scope exit
class DeviceTables {
   public:
    DeviceTables(uint32_t n_vars, const std::vector<uint32_t>& degree, const std::vector<const uint64_t*>& tables) {
        Context::check(zksc_tables_upload(Context::get(), n_vars, 1, (uint32_t)degree.size(), degree.data(), tables.data(), &h_));
    }
    ~DeviceTables() { zksc_tables_free(h_); }
    DeviceTables(const DeviceTables&) = delete;
    DeviceTables& operator=(const DeviceTables&) = delete;
    zksc_tables* h() const { return h_; }

   private:
    zksc_tables* h_ = nullptr;
};

// ---- polynomial/src/multilinear/evaluation_form.rs -------------------------------------------------------------------
class Multilinear {
   public:
    size_t n_vars = 0;
    std::vector<Fr> evaluations;

    Multilinear() = default;
    static Multilinear new_(std::vector<Fr> evaluations) {  // :12-26
        const size_t n = evaluations.size();
        if (n == 0 || (n & (n - 1))) throw Error(ZKSC_ERR_SHAPE, "Number of evaluations must be a power of 2");
        Multilinear m;
        while ((size_t(1) << m.n_vars) < n) m.n_vars++;
        m.evaluations = std::move(evaluations);
        return m;
    }
    bool operator==(const Multilinear& o) const { return n_vars == o.n_vars && evaluations == o.evaluations; }

    Multilinear add_distinct(const Multilinear& rhs) const { return outer(rhs, 0); }  // :28-39
    Multilinear mul_distinct(const Multilinear& rhs) const { return outer(rhs, 1); }  // :41-52

    std::vector<uint8_t> to_bytes() const {  // :54-62
        DeviceTables t((uint32_t)n_vars, {1}, {raw(evaluations)});
        std::vector<uint8_t> out(evaluations.size() * 32);
        Context::check(zksc_tables_to_bytes(t.h(), 0, out.data()));
        return out;
    }
    Multilinear split_poly_into_two_and_sum_each_part() const {  // :68-74
        if (n_vars == 0) throw Error(ZKSC_ERR_SHAPE, "a constant has no halves");
        DeviceTables t((uint32_t)n_vars, {1}, {raw(evaluations)});
        std::vector<Fr> h(2);
        Context::check(zksc_round_evals(t.h(), raw(h)));
        return new_(h);
    }
    Fr sum_over_the_boolean_hypercube() const {  // :80-84
        DeviceTables t((uint32_t)n_vars, {1}, {raw(evaluations)});
        Fr s;
        Context::check(zksc_poly_sum(t.h(), s.v));
        return s;
    }
    Multilinear partial_evaluation(const Fr& eval_point, size_t variable_index) const {  // :123-141
        std::vector<Fr> out(evaluations.size() / 2 ? evaluations.size() / 2 : 1);
        Context::check(zksc_ml_partial_evaluation(Context::get(), raw(evaluations), evaluations.size(), eval_point.v, (uint32_t)variable_index, raw(out)));
        return new_(out);
    }
    Multilinear partial_evaluations(const std::vector<Fr>& points, const std::vector<size_t>& variable_indices) const {  // :143-159
        if (points.size() != variable_indices.size()) throw Error(ZKSC_ERR_SHAPE, "The length of evaluation_points and variable_indices should be the same");
        Multilinear e = *this;
        for (size_t i = 0; i < points.size(); i++) e = e.partial_evaluation(points[i], variable_indices[i]);
        return e;
    }
    Fr evaluation(const std::vector<Fr>& evaluation_points) const {  // :162-175
        if (evaluation_points.size() != n_vars) throw Error(ZKSC_ERR_SHAPE, "Number of evaluation points must match the number of variables");
        Fr out;
        Context::check(zksc_ml_evaluation(Context::get(), raw(evaluations), evaluations.size(), raw(evaluation_points), (uint32_t)evaluation_points.size(), out.v));
        return out;
    }
    Multilinear operator+(const Multilinear& rhs) const { return ew(0, rhs.evaluations); }  // impl Add :178-194
    Multilinear operator-(const Multilinear& rhs) const { return ew(1, rhs.evaluations); }  // impl Sub :209-225
    Multilinear operator*(const Fr& scalar) const { return ew(3, std::vector<Fr>{scalar}); }  // impl Mul<F> :235-251
    Multilinear element_mul(const Multilinear& rhs) const { return ew(2, rhs.evaluations); }

   private:
    Multilinear outer(const Multilinear& rhs, int mul) const {
        std::vector<Fr> out(evaluations.size() * rhs.evaluations.size());
        Context::check(zksc_ml_outer(Context::get(), mul, raw(evaluations), evaluations.size(), raw(rhs.evaluations), rhs.evaluations.size(), raw(out)));
        return new_(out);
    }
    Multilinear ew(int op, const std::vector<Fr>& other) const {
        if (op != 3 && other.size() != evaluations.size()) throw Error(ZKSC_ERR_SHAPE, "operands must have the same number of variables");
        std::vector<Fr> out(evaluations.size());
        Context::check(zksc_ml_elementwise(Context::get(), op, raw(evaluations), raw(other), evaluations.size(), raw(out)));
        return new_(out);
    }
};

// ---- polynomial/src/composed/composed_multilinear.rs ---------------------------------------------------------------
class ComposedMultilinear {
   public:
    std::vector<Multilinear> polys;

    static ComposedMultilinear new_(std::vector<Multilinear> polys) {  // :12-18
        if (polys.empty()) throw Error(ZKSC_ERR_SHAPE, "a product needs at least one factor");
        for (const auto& p : polys)
            if (p.n_vars != polys[0].n_vars) throw Error(ZKSC_ERR_SHAPE, "all factors must have the same number of variables");
        ComposedMultilinear c;
        c.polys = std::move(polys);
        return c;
    }
    size_t n_vars() const { return polys[0].n_vars; }   // :20-22
    size_t max_degree() const { return polys.size(); }  // :101-103
    std::vector<uint8_t> to_bytes() const {             // :40-48
        std::vector<uint8_t> out;
        for (const auto& p : polys) { auto b = p.to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
        return out;
    }
    ComposedMultilinear partial_evaluation(const Fr& point, size_t variable_index) const {  // :63-75
        std::vector<Multilinear> out;
        for (const auto& p : polys) out.push_back(p.partial_evaluation(point, variable_index));
        return new_(out);
    }
    Fr evaluation(const std::vector<Fr>& points) const {  // :52-61
        Fr r = Fr::one();
        for (const auto& p : polys) r = r * p.evaluation(points);
        return r;
    }
    std::vector<Fr> element_wise_product() const {  // :105-111
        Multilinear acc = polys[0];
        for (size_t i = 1; i < polys.size(); i++) acc = acc.element_mul(polys[i]);
        return acc.evaluations;
    }
    std::vector<Fr> element_wise_add() const {  // :113-119
        Multilinear acc = polys[0];
        for (size_t i = 1; i < polys.size(); i++) acc = acc + polys[i];
        return acc.evaluations;
    }
};

inline std::unique_ptr<DeviceTables> upload(const std::vector<ComposedMultilinear>& poly) {
    if (poly.empty()) throw Error(ZKSC_ERR_SHAPE, "no polynomial");
    std::vector<uint32_t> deg;
    std::vector<const uint64_t*> tabs;
    for (const auto& p : poly) {
        if (p.n_vars() != poly[0].n_vars()) throw Error(ZKSC_ERR_SHAPE, "all products must have the same number of variables");
        deg.push_back((uint32_t)p.max_degree());
        for (const auto& m : p.polys) tabs.push_back(raw(m.evaluations));
    }
    return std::make_unique<DeviceTables>((uint32_t)poly[0].n_vars(), deg, tabs);
}

// ---- transcripts/fiat-shamir/src/fiat_shamir.rs --------------------------------------------------------------------
class FiatShamirTranscript {
   public:
    FiatShamirTranscript() : h_(zksc_transcript_new()) {}
    static FiatShamirTranscript new_() { return FiatShamirTranscript(); }
    FiatShamirTranscript(FiatShamirTranscript&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    FiatShamirTranscript(const FiatShamirTranscript&) = delete;
    ~FiatShamirTranscript() { if (h_) zksc_transcript_free(h_); }
    void commit(const std::vector<uint8_t>& bytes) { zksc_transcript_commit(h_, bytes.data(), bytes.size()); }  // :17-19
    std::array<uint8_t, 32> challenge() { std::array<uint8_t, 32> d; zksc_transcript_challenge(h_, d.data()); return d; }  // :21-25
    Fr evaluate_challenge_into_field() { Fr r; zksc_transcript_challenge_field(h_, r.v); return r; }  // :27-29
    std::vector<Fr> evaluate_n_challenge_into_field(size_t n) {  // :31-39
        std::vector<Fr> r;
        for (size_t i = 0; i < n; i++) r.push_back(evaluate_challenge_into_field());
        return r;
    }

   private:
    zksc_transcript* h_;
};

// ---- polynomial/src/univariate/sparse_univariate.rs ------------------------------------------------------------------
struct UnivariateMonomial {
    Fr coeff, pow;
    bool operator==(const UnivariateMonomial& o) const { return coeff == o.coeff && pow == o.pow; }
};
static_assert(sizeof(UnivariateMonomial) == 64, "(coeff, pow) pairs are passed to the library as they are");
struct SparseUnivariatePolynomial {
    std::vector<UnivariateMonomial> monomial;
    static SparseUnivariatePolynomial zero() { return {}; }  // :23-25
    // interpolation over x = 0 .. ys.len()-1, the only form the sumcheck path uses (:40-63)
    static SparseUnivariatePolynomial interpolation_evals(const std::vector<Fr>& ys) {
        SparseUnivariatePolynomial p;
        p.monomial.resize(ys.size());
        p.monomial.resize(zksc_sparse_interpolate(raw(ys), (uint32_t)ys.size(), reinterpret_cast<uint64_t*>(p.monomial.data())));
        return p;
    }
    SparseUnivariatePolynomial operator+(const SparseUnivariatePolynomial& rhs) const {  // impl Add :159-203
        SparseUnivariatePolynomial p;
        p.monomial.resize(monomial.size() + rhs.monomial.size() + 1);
        p.monomial.resize(zksc_sparse_add(reinterpret_cast<const uint64_t*>(monomial.data()), (uint32_t)monomial.size(),
                                          reinterpret_cast<const uint64_t*>(rhs.monomial.data()), (uint32_t)rhs.monomial.size(),
                                          reinterpret_cast<uint64_t*>(p.monomial.data())));
        return p;
    }
    Fr evaluate(const Fr& point) const {  // :90-106
        Fr r;
        zksc_sparse_evaluate(reinterpret_cast<const uint64_t*>(monomial.data()), (uint32_t)monomial.size(), point.v, r.v);
        return r;
    }
    std::vector<uint8_t> to_bytes() const {  // :27-34
        std::vector<uint8_t> out;
        for (const auto& m : monomial) { append(out, m.coeff); append(out, m.pow); }
        return out;
    }
};

namespace detail {
struct Rounds {  // the wire form of zksc_prove / zksc_verify_rounds
    uint32_t n = 0, stride = 0;
    std::vector<Fr> msgs, challenges;
    std::vector<uint32_t> lens;
};
inline Rounds prove(zksc_tables* t, int protocol, uint32_t n_vars, const std::vector<uint32_t>& deg, const Fr* sum) {
    Rounds r;
    r.n = n_vars;
    r.stride = zksc_msg_stride(protocol, (uint32_t)deg.size(), deg.data());
    r.msgs.resize((size_t)n_vars * r.stride + 1);
    r.lens.resize(n_vars + 1);
    r.challenges.resize(n_vars + 1);
    Context::check(zksc_prove(t, protocol, sum ? sum->v : nullptr, raw(r.msgs), r.lens.data(), raw(r.challenges)));
    r.challenges.resize(n_vars);
    return r;
}
// Ok((sub-claim sum, challenges)) or Err("Verification failed")
inline Result<std::pair<Fr, std::vector<Fr>>> verify_rounds(int protocol, const Rounds& r, const Fr& sum, const std::vector<uint8_t>& prefix = {}) {
    Fr sub;
    std::vector<Fr> ch(r.n + 1);
    std::vector<Fr> msgs = r.msgs;
    msgs.resize(msgs.size() + 1);
    std::vector<uint32_t> lens = r.lens;
    lens.resize(lens.size() + 1);
    int rc = zksc_verify_rounds(protocol, r.n, r.stride, sum.v, raw(msgs), lens.data(), prefix.empty() ? nullptr : prefix.data(), prefix.size(), sub.v, raw(ch));
    if (rc == ZKSC_ERR_VERIFY) return {std::nullopt, "Verification failed"};
    Context::check(rc);
    ch.resize(r.n);
    return {std::make_pair(sub, ch), nullptr};
}
}  // namespace detail

// ---- sumcheck/src/sumcheck.rs ----------------------------------------------------------------------------------------
struct SumcheckProof {  // :11-15
    Multilinear poly;
    Fr sum;
    std::vector<Multilinear> univariate_poly;
};
class Sumcheck {
   public:
    Multilinear poly;
    Fr sum;
    static Sumcheck new_(Multilinear poly) { Sumcheck s; s.poly = std::move(poly); return s; }  // :18-23
    void poly_sum() { sum = poly.sum_over_the_boolean_hypercube(); }                           // :25-27
    std::pair<SumcheckProof, std::vector<Fr>> prove() const {                                   // :29-61
        DeviceTables t((uint32_t)poly.n_vars, {1}, {raw(poly.evaluations)});
        detail::Rounds r = detail::prove(t.h(), ZKSC_PROTO_SUMCHECK, (uint32_t)poly.n_vars, {1}, &sum);
        SumcheckProof p{poly, sum, {}};
        for (uint32_t j = 0; j < r.n; j++) p.univariate_poly.push_back(Multilinear::new_({r.msgs[2 * j], r.msgs[2 * j + 1]}));
        return {p, r.challenges};
    }
    bool verify(const SumcheckProof& proof) const {  // :63-95
        detail::Rounds r;
        r.n = (uint32_t)proof.univariate_poly.size();
        r.stride = 2;
        for (const auto& u : proof.univariate_poly) {
            if (u.evaluations.size() != 2) return false;
            r.msgs.push_back(u.evaluations[0]);
            r.msgs.push_back(u.evaluations[1]);
            r.lens.push_back(2);
        }
        auto v = detail::verify_rounds(ZKSC_PROTO_SUMCHECK, r, proof.sum);
        if (!v.is_ok()) return false;
        return proof.poly.evaluation(v.unwrap().second) == v.unwrap().first;
    }
};

// ---- sumcheck/src/composed/composed_sumcheck.rs ----------------------------------------------------------------------
namespace composed {
struct ComposedSumcheckProof {  // :15-18
    ComposedMultilinear poly;
    std::vector<std::vector<Fr>> round_polys;
};
}  // namespace composed
class ComposedSumcheck {
   public:
    ComposedMultilinear poly;
    Fr sum;
    static ComposedSumcheck new_(ComposedMultilinear poly) { ComposedSumcheck s; s.poly = std::move(poly); return s; }  // :21-26
    static Fr calculate_poly_sum(const ComposedMultilinear& poly) {                                                     // :28-30
        auto t = upload({poly});
        Fr s;
        Context::check(zksc_poly_sum(t->h(), s.v));
        return s;
    }
    std::pair<composed::ComposedSumcheckProof, std::vector<Fr>> prove() const {  // :32-67
        auto t = upload({poly});
        const uint32_t d = (uint32_t)poly.max_degree();
        detail::Rounds r = detail::prove(t->h(), ZKSC_PROTO_COMPOSED, (uint32_t)poly.n_vars(), {d}, nullptr);
        composed::ComposedSumcheckProof p{poly, {}};
        for (uint32_t j = 0; j < r.n; j++) p.round_polys.emplace_back(r.msgs.begin() + (size_t)j * r.stride, r.msgs.begin() + (size_t)j * r.stride + d + 1);
        return {p, r.challenges};
    }
    bool verify(const composed::ComposedSumcheckProof& proof, const Fr& sum_) const {  // :69-95
        detail::Rounds r;
        r.n = (uint32_t)proof.round_polys.size();
        for (const auto& rp : proof.round_polys) r.stride = rp.size() > r.stride ? (uint32_t)rp.size() : r.stride;
        for (const auto& rp : proof.round_polys) {
            r.lens.push_back((uint32_t)rp.size());
            for (uint32_t i = 0; i < r.stride; i++) r.msgs.push_back(i < rp.size() ? rp[i] : Fr::zero());
        }
        auto v = detail::verify_rounds(ZKSC_PROTO_COMPOSED, r, sum_);
        if (!v.is_ok()) return false;
        return proof.poly.evaluation(v.unwrap().second) == v.unwrap().first;
    }
};

// ---- sumcheck/src/composed/multi_composed_sumcheck.rs ----------------------------------------------------------------
struct ComposedSumcheckProof {  // :12-16
    std::vector<SparseUnivariatePolynomial> round_polys;
    Fr sum;
    std::vector<uint8_t> to_bytes() const {  // :24-32
        std::vector<uint8_t> out;
        for (const auto& rp : round_polys) { auto b = rp.to_bytes(); out.insert(out.end(), b.begin(), b.end()); }
        return out;
    }
};
struct SubClaim {  // :18-22
    Fr sum;
    std::vector<Fr> challenges;
};
namespace detail {
inline ComposedSumcheckProof proof_of(const Rounds& r, const Fr& sum) {
    ComposedSumcheckProof p;
    p.sum = sum;
    for (uint32_t j = 0; j < r.n; j++) {
        SparseUnivariatePolynomial rp;
        for (uint32_t m = 0; m < r.lens[j]; m++) rp.monomial.push_back({r.msgs[(size_t)j * r.stride + 2 * m], r.msgs[(size_t)j * r.stride + 2 * m + 1]});
        p.round_polys.push_back(rp);
    }
    return p;
}
inline Rounds rounds_of(const ComposedSumcheckProof& proof) {
    Rounds r;
    r.n = (uint32_t)proof.round_polys.size();
    size_t mx = 1;
    for (const auto& rp : proof.round_polys) mx = rp.monomial.size() > mx ? rp.monomial.size() : mx;
    r.stride = (uint32_t)(2 * mx);
    for (const auto& rp : proof.round_polys) {
        r.lens.push_back((uint32_t)rp.monomial.size());
        for (size_t m = 0; m < mx; m++) {
            r.msgs.push_back(m < rp.monomial.size() ? rp.monomial[m].coeff : Fr::zero());
            r.msgs.push_back(m < rp.monomial.size() ? rp.monomial[m].pow : Fr::zero());
        }
    }
    return r;
}
}  // namespace detail
struct MultiComposedSumcheckProver {
    static Fr calculate_poly_sum(const std::vector<ComposedMultilinear>& poly) {  // :37-45
        auto t = upload(poly);
        Fr s;
        Context::check(zksc_poly_sum(t->h(), s.v));
        return s;
    }
    using Proved = std::pair<ComposedSumcheckProof, std::vector<Fr>>;
    static Result<Proved> prove(const std::vector<ComposedMultilinear>& poly, const Fr& sum) { return run(poly, sum, ZKSC_PROTO_MULTI_FULL); }             // :47-54
    static Result<Proved> prove_partial(const std::vector<ComposedMultilinear>& poly, const Fr& sum) { return run(poly, sum, ZKSC_PROTO_MULTI_PARTIAL); }  // :56-62

   private:
    static Result<Proved> run(const std::vector<ComposedMultilinear>& poly, const Fr& sum, int protocol) {  // prove_internal :64-120
        auto t = upload(poly);
        std::vector<uint32_t> deg;
        for (const auto& p : poly) deg.push_back((uint32_t)p.max_degree());
        detail::Rounds r = detail::prove(t->h(), protocol, (uint32_t)poly[0].n_vars(), deg, &sum);
        return {Proved{detail::proof_of(r, sum), r.challenges}, nullptr};
    }
};
struct MultiComposedSumcheckVerifier {
    static Result<bool> verify(const std::vector<ComposedMultilinear>& poly, const ComposedSumcheckProof& proof) {  // :126-142
        std::vector<uint8_t> prefix;  // composed_poly_to_bytes (sumcheck/src/utils.rs:53-59)
        for (const auto& p : poly) { auto b = p.to_bytes(); prefix.insert(prefix.end(), b.begin(), b.end()); }
        auto v = detail::verify_rounds(ZKSC_PROTO_MULTI_FULL, detail::rounds_of(proof), proof.sum, prefix);
        if (!v.is_ok()) return {std::nullopt, v.err};
        auto t = upload(poly);
        Fr val;
        const auto& ch = v.unwrap().second;
        if (ch.empty()) Context::check(zksc_poly_sum(t->h(), val.v));
        else Context::check(zksc_evaluate(t->h(), raw(ch), val.v));   // the oracle check, n folds on the device
        return {val == v.unwrap().first, nullptr};
    }
    static Result<SubClaim> verify_partial(const ComposedSumcheckProof& proof) {  // :143-149
        auto v = detail::verify_rounds(ZKSC_PROTO_MULTI_PARTIAL, detail::rounds_of(proof), proof.sum);
        if (!v.is_ok()) return {std::nullopt, v.err};
        return {SubClaim{v.unwrap().first, v.unwrap().second}, nullptr};
    }
};

// ---- circuit/src/{gate,circuit,utils}.rs, gkr/src/protocol.rs --------------------------------------------------------
enum class GateType { Add, Mul };  // gate.rs:2-5
struct Gate {                      // gate.rs:8-17
    GateType gate_type;
    std::array<size_t, 2> inputs;
    static Gate new_(GateType t, std::array<size_t, 2> in) { return Gate{t, in}; }
};
struct CircuitLayer {  // circuit.rs:10-25
    std::vector<Gate> layer;
    static CircuitLayer new_(std::vector<Gate> l) { return CircuitLayer{std::move(l)}; }
};
struct Circuit {  // circuit.rs:15-122
    std::vector<CircuitLayer> layers;
    static Circuit new_(std::vector<CircuitLayer> l) { return Circuit{std::move(l)}; }
    std::vector<std::vector<Fr>> evaluation(const std::vector<Fr>& input) const {  // :32-55 (one pass over the gates: the caller's work, host)
        std::vector<std::vector<Fr>> out{input};
        for (size_t li = layers.size(); li-- > 0;) {
            const std::vector<Fr>& cur = out.back();
            std::vector<Fr> next;
            for (const Gate& g : layers[li].layer) {
                if (g.inputs[0] >= cur.size() || g.inputs[1] >= cur.size()) throw Error(ZKSC_ERR_SHAPE, "gate input index out of range");
                next.push_back(g.gate_type == GateType::Add ? cur[g.inputs[0]] + cur[g.inputs[1]] : cur[g.inputs[0]] * cur[g.inputs[1]]);
            }
            out.push_back(next);
        }
        return {out.rbegin(), out.rend()};
    }
    // :57-95 with circuit/src/utils.rs:1-34 -- the dense 0/1 wiring tables (the prover never materialises them beyond layer 0)
    std::pair<Multilinear, Multilinear> add_mult_mle(size_t layer_index) const {
        const size_t bits = layer_index + 1, n = layer_index == 0 ? 8 : size_t(1) << (layer_index + 2 * bits);
        std::vector<Fr> add(n), mul(n);
        const auto& gates = layers.at(layer_index).layer;
        for (size_t gi = 0; gi < gates.size(); gi++) {
            const Gate& g = gates[gi];
            if ((gi >> (layer_index ? layer_index : 1)) || (g.inputs[0] >> bits) || (g.inputs[1] >> bits)) throw Error(ZKSC_ERR_SHAPE, "gate label does not fit its bit field");
            (g.gate_type == GateType::Add ? add : mul)[(gi << (2 * bits)) | (g.inputs[0] << bits) | g.inputs[1]] = Fr::one();
        }
        return {Multilinear::new_(add), Multilinear::new_(mul)};
    }
    static Circuit random(size_t num_of_layers) {  // :97-121
        Circuit c;
        for (size_t li = 0; li < num_of_layers; li++) {
            const size_t n_in = size_t(2) << li;
            CircuitLayer l;
            for (size_t g = 0; g < (size_t(1) << li); g++) l.layer.push_back(Gate{li % 2 == 0 ? GateType::Add : GateType::Mul, {(g * 2) % n_in, (g * 2 + 1) % n_in}});
            c.layers.push_back(l);
        }
        return c;
    }
};
struct GKRProof {  // protocol.rs:10-15
    std::vector<ComposedSumcheckProof> sumcheck_proofs;
    std::vector<Fr> wb_s, wc_s;
    Multilinear w_0_mle;
};
struct GKRProtocol {
    // protocol.rs:21-113: the whole proof is ONE library call (zksc_gkr_prove)
    static GKRProof prove(const Circuit& circuit, const std::vector<std::vector<Fr>>& circuit_evaluation) {
        const uint32_t L = (uint32_t)circuit.layers.size();
        if (circuit_evaluation.size() != (size_t)L + 1) throw Error(ZKSC_ERR_SHAPE, "circuit_evaluation must hold one vector per layer plus the input");
        std::vector<uint32_t> n_gates, in0, in1;
        std::vector<uint8_t> type;
        for (const auto& l : circuit.layers) {
            n_gates.push_back((uint32_t)l.layer.size());
            for (const Gate& g : l.layer) {
                type.push_back(g.gate_type == GateType::Add ? 0 : 1);
                in0.push_back((uint32_t)g.inputs[0]);
                in1.push_back((uint32_t)g.inputs[1]);
            }
        }
        std::vector<const uint64_t*> vals;
        std::vector<uint64_t> vlen;
        for (const auto& v : circuit_evaluation) { vals.push_back(raw(v)); vlen.push_back(v.size()); }
        const size_t rounds = zksc_gkr_total_rounds(L);
        std::vector<Fr> w0(2), sums(L), wb(L), wc(L), msgs(rounds * 6 + 1), chal(rounds + 1);
        std::vector<uint32_t> lens(rounds + 1);
        Context::check(zksc_gkr_prove(Context::get(), L, n_gates.data(), type.data(), in0.data(), in1.data(), vals.data(), vlen.data(), raw(w0), raw(sums), raw(wb),
                                      raw(wc), raw(msgs), lens.data(), raw(chal)));
        GKRProof proof;
        proof.w_0_mle = Multilinear::new_(w0);
        proof.wb_s = wb;
        proof.wc_s = wc;
        size_t off = 0;
        for (uint32_t li = 0; li < L; li++) {
            detail::Rounds r;
            r.n = 2 * (li + 1);
            r.stride = 6;
            r.msgs.assign(msgs.begin() + off * 6, msgs.begin() + (off + r.n) * 6);
            r.lens.assign(lens.begin() + off, lens.begin() + off + r.n);
            proof.sumcheck_proofs.push_back(detail::proof_of(r, sums[li]));
            off += r.n;
        }
        return proof;
    }
    // protocol.rs:115-195 (+ generate_layer_one_verify_sumcheck, gkr/src/utils.rs:59-98).  Where the reference unwrap()s a failed
    // verify_partial (a panic) this returns false.
    static bool verify(const Circuit& circuit, const std::vector<Fr>& input, const GKRProof& proof) {
        if (proof.sumcheck_proofs.size() != proof.wb_s.size() || proof.sumcheck_proofs.size() != proof.wc_s.size() || proof.sumcheck_proofs.empty()) return false;
        FiatShamirTranscript transcript;
        transcript.commit(proof.w_0_mle.to_bytes());
        std::vector<Fr> n_r = transcript.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars);
        Fr claimed_sum = proof.w_0_mle.evaluation(n_r);
        std::vector<Fr> r_b, r_c;
        Fr alpha = Fr::zero(), beta = Fr::zero();
        {
            const ComposedSumcheckProof& p0 = proof.sumcheck_proofs[0];
            if (claimed_sum != p0.sum) return false;
            transcript.commit(p0.to_bytes());
            auto sub = MultiComposedSumcheckVerifier::verify_partial(p0);
            if (!sub.is_ok()) return false;
            auto [add_mle_1, mult_mle_1] = circuit.add_mult_mle(0);
            std::vector<Fr> rbc = n_r;
            rbc.insert(rbc.end(), sub.unwrap().challenges.begin(), sub.unwrap().challenges.end());
            const Fr wb = proof.wb_s[0], wc = proof.wc_s[0];
            if (add_mle_1.evaluation(rbc) * (wb + wc) + mult_mle_1.evaluation(rbc) * (wb * wc) != sub.unwrap().sum) return false;
            const Fr a = transcript.evaluate_challenge_into_field(), b = transcript.evaluate_challenge_into_field();
            claimed_sum = a * wb + b * wc;
        }
        for (size_t i = 1; i < proof.sumcheck_proofs.size(); i++) {
            const ComposedSumcheckProof& p = proof.sumcheck_proofs[i];
            if (claimed_sum != p.sum) return false;
            transcript.commit(p.to_bytes());
            auto sub = MultiComposedSumcheckVerifier::verify_partial(p);
            if (!sub.is_ok()) return false;
            const auto& ch = sub.unwrap().challenges;
            r_b.assign(ch.begin(), ch.begin() + ch.size() / 2);
            r_c.assign(ch.begin() + ch.size() / 2, ch.end());
            alpha = transcript.evaluate_challenge_into_field();
            beta = transcript.evaluate_challenge_into_field();
            claimed_sum = alpha * proof.wb_s[i] + beta * proof.wc_s[i];
        }
        const Multilinear w_mle_input = Multilinear::new_(input);
        return claimed_sum == alpha * w_mle_input.evaluation(r_b) + beta * w_mle_input.evaluation(r_c);
    }
};

// ---- layered circuits of any power-of-two widths, proved in time linear in their gates (zksc_circuit_*, zksc_gkr_prove_linear) ----
// BASELINE config 4 as written ("width 2^20, depth 8").  `Circuit` above, like the reference's, holds only the pyramid shape; on a
// pyramid `LayeredCircuit::prove` returns the GKRProof of GKRProtocol::prove, byte for byte, and `verify` is GKRProtocol::verify
// (protocol.rs:115-195) with the wiring polynomials taken from the gate lists on the device (zksc_circuit_wiring_eval).
class LayeredCircuit {
   public:
    // log_width[i] = log2(gates of layer i), output layer first, last entry = log2(inputs); layers[i] = the gates of layer i
    LayeredCircuit(std::vector<uint32_t> log_width, const std::vector<CircuitLayer>& layers) : lw_(std::move(log_width)) {
        if (lw_.size() != layers.size() + 1) throw Error(ZKSC_ERR_SHAPE, "one width per layer plus the input layer");
        std::vector<uint32_t> in0, in1;
        std::vector<uint8_t> type;
        for (size_t i = 0; i < layers.size(); i++) {
            if (layers[i].layer.size() != (size_t(1) << lw_[i])) throw Error(ZKSC_ERR_SHAPE, "a layer must have 2^log_width gates");
            for (const Gate& g : layers[i].layer) {
                type.push_back(g.gate_type == GateType::Add ? 0 : 1);
                in0.push_back((uint32_t)g.inputs[0]);
                in1.push_back((uint32_t)g.inputs[1]);
            }
        }
        Context::check(zksc_circuit_create(Context::get(), (uint32_t)layers.size(), lw_.data(), type.data(), in0.data(), in1.data(), &h_));
    }
    static LayeredCircuit from_circuit(const Circuit& c) {
        std::vector<uint32_t> lw;
        for (uint32_t i = 0; i <= c.layers.size(); i++) lw.push_back(i);
        return LayeredCircuit(lw, c.layers);
    }
    LayeredCircuit(const LayeredCircuit&) = delete;
    LayeredCircuit& operator=(const LayeredCircuit&) = delete;
    LayeredCircuit(LayeredCircuit&& o) noexcept : lw_(std::move(o.lw_)), h_(o.h_) { o.h_ = nullptr; }
    ~LayeredCircuit() { if (h_) zksc_circuit_free(h_); }

    // Circuit::evaluation (circuit.rs:32-55) on the device: returns the output layer, every layer stays in HBM for prove()
    std::vector<Fr> evaluate(const std::vector<Fr>& input) {
        if (input.size() != (size_t(1) << lw_.back())) throw Error(ZKSC_ERR_SHAPE, "the input layer has 2^log_width values");
        std::vector<Fr> out(size_t(1) << lw_[0]);
        Context::check(zksc_circuit_evaluate(h_, raw(input), raw(out)));
        return out;
    }
    GKRProof prove() {
        const uint32_t L = (uint32_t)lw_.size() - 1;
        const size_t rounds = zksc_circuit_total_rounds(h_);
        std::vector<Fr> w0(std::max<size_t>(2, size_t(1) << lw_[0])), sums(L), wb(L), wc(L), msgs(rounds * 6 + 1), chal(rounds + 1);
        std::vector<uint32_t> lens(rounds + 1);
        Context::check(zksc_gkr_prove_linear(h_, raw(w0), raw(sums), raw(wb), raw(wc), raw(msgs), lens.data(), raw(chal)));
        GKRProof proof;
        proof.w_0_mle = Multilinear::new_(w0);
        proof.wb_s = wb;
        proof.wc_s = wc;
        size_t off = 0;
        for (uint32_t li = 0; li < L; li++) {
            detail::Rounds r;
            r.n = 2 * lw_[li + 1];
            r.stride = 6;
            r.msgs.assign(msgs.begin() + off * 6, msgs.begin() + (off + r.n) * 6);
            r.lens.assign(lens.begin() + off, lens.begin() + off + r.n);
            proof.sumcheck_proofs.push_back(detail::proof_of(r, sums[li]));
            off += r.n;
        }
        return proof;
    }
    bool verify(const std::vector<Fr>& input, const GKRProof& proof) {
        const size_t L = lw_.size() - 1;
        if (proof.sumcheck_proofs.size() != L || proof.wb_s.size() != L || proof.wc_s.size() != L) return false;
        FiatShamirTranscript transcript;
        transcript.commit(proof.w_0_mle.to_bytes());
        std::vector<Fr> r_b = transcript.evaluate_n_challenge_into_field(proof.w_0_mle.n_vars), r_c;
        Fr claimed_sum = proof.w_0_mle.evaluation(r_b);
        Fr alpha = Fr::one(), beta = Fr::zero();
        for (size_t i = 0; i < L; i++) {
            const ComposedSumcheckProof& p = proof.sumcheck_proofs[i];
            if (claimed_sum != p.sum) return false;
            transcript.commit(p.to_bytes());
            auto sub = MultiComposedSumcheckVerifier::verify_partial(p);
            if (!sub.is_ok()) return false;
            const auto& ch = sub.unwrap().challenges;
            if (ch.size() != 2 * (size_t)lw_[i + 1]) return false;
            const std::vector<Fr> b(ch.begin(), ch.begin() + ch.size() / 2), c(ch.begin() + ch.size() / 2, ch.end());
            Fr wiring[2];
            Context::check(zksc_circuit_wiring_eval(h_, (uint32_t)i, raw(r_b), alpha.v, i ? raw(r_c) : nullptr, i ? beta.v : nullptr, raw(b), raw(c), wiring[0].v));
            const Fr wb = proof.wb_s[i], wc = proof.wc_s[i];
            if (wiring[0] * (wb + wc) + wiring[1] * (wb * wc) != sub.unwrap().sum) return false;
            alpha = transcript.evaluate_challenge_into_field();
            beta = transcript.evaluate_challenge_into_field();
            claimed_sum = alpha * wb + beta * wc;
            r_b = b;
            r_c = c;
        }
        const Multilinear w_mle_input = Multilinear::new_(input);
        return claimed_sum == alpha * w_mle_input.evaluation(r_b) + beta * w_mle_input.evaluation(r_c);
    }

   private:
    std::vector<uint32_t> lw_;
    zksc_circuit* h_ = nullptr;
};

}  // namespace zk
#endif  // ZKSC_HPP
