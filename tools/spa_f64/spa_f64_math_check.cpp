// Host check of qkd_ldpc_v_b200/csrc/spa_f64_math.hpp against glibc: g++ -O2 -std=c++17 -o /tmp/spa_chk tools/spa_f64/spa_f64_math_check.cpp && /tmp/spa_chk
// (-ffp-contract=off is not needed: the header calls fma explicitly.) Prints the largest error in ulp per argument range.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../qkd_ldpc_v_b200/csrc/spa_f64_math.hpp"

static double ulp_err(double got, double want) {
    if (got == want) return 0;
    if (std::isnan(got) && std::isnan(want)) return 0;
    if (std::isinf(want) || std::isinf(got) || std::isnan(got) || std::isnan(want)) return 1e300;
    const double u = std::ldexp(1.0, std::ilogb(want) - 52);
    return std::fabs(got - want) / u;
}

int main() {
    std::mt19937_64 g(12345);
    std::uniform_real_distribution<double> U(0, 1);
    const long N = 10000000;
    int bad = 0;
    // tanh(x/2) by magnitude decade
    const double tl[] = {1e-300, 1e-10, 1e-3, 0.1, 0.6, 1.0, 5.0, 30.0, 36.0, 39.0, 100.0, 1e5};
    for (size_t i = 0; i + 1 < sizeof tl / sizeof *tl; ++i) {
        double worst = 0, wx = 0;
        for (long k = 0; k < N / 4; ++k) {
            const double x = (tl[i] + (tl[i + 1] - tl[i]) * U(g)) * ((k & 1) ? -1 : 1);
            const double e = ulp_err(qk::spa_tanh_half_f64(x), std::tanh(x / 2));
            if (e > worst) { worst = e; wx = x; }
        }
        printf("tanh(x/2)  |x| in [%g, %g): max %.2f ulp at %.17g\n", tl[i], tl[i + 1], worst, wx);
        bad += worst > 5.5;
    }
    const double al[] = {0, 1e-8, 1e-3, 0.1, 0.17, 0.2, 0.5, 0.9, 0.999, 0.999999, 1.0};
    for (size_t i = 0; i + 1 < sizeof al / sizeof *al; ++i) {
        double worst = 0, wy = 0;
        for (long k = 0; k < N / 4; ++k) {
            const double y = (al[i] + (al[i + 1] - al[i]) * U(g)) * ((k & 1) ? -1 : 1);
            const double e = ulp_err(qk::spa_two_atanh_f64(y), 2 * std::atanh(y));
            if (e > worst) { worst = e; wy = y; }
        }
        printf("2atanh(y)  |y| in [%g, %g): max %.2f ulp at %.17g\n", al[i], al[i + 1], worst, wy);
        bad += worst > 5.5;
    }
    // near the pole: y = 1 - j * 2^-53
    {
        double worst = 0;
        for (int j = 1; j < 100000; ++j) {
            const double y = 1.0 - j * std::ldexp(1.0, -53);
            worst = std::fmax(worst, ulp_err(qk::spa_two_atanh_f64(y), 2 * std::atanh(y)));
            worst = std::fmax(worst, ulp_err(qk::spa_two_atanh_f64(-y), 2 * std::atanh(-y)));
        }
        printf("2atanh(1 - j 2^-53), j < 1e5: max %.2f ulp\n", worst);
        bad += worst > 5.5;
    }
    // special values
    const double inf = INFINITY, nan = NAN;
    struct { double got, want; const char *what; } sp[] = {
        {qk::spa_tanh_half_f64(0.0), 0.0, "tanh(+0)"}, {qk::spa_tanh_half_f64(-0.0), -0.0, "tanh(-0)"},
        {qk::spa_tanh_half_f64(inf), 1.0, "tanh(inf)"}, {qk::spa_tanh_half_f64(-inf), -1.0, "tanh(-inf)"},
        {qk::spa_tanh_half_f64(nan), nan, "tanh(nan)"}, {qk::spa_tanh_half_f64(200.0), 1.0, "tanh(100)"},
        {qk::spa_tanh_half_f64(1e-320), std::tanh(0.5e-320), "tanh(denormal)"},
        {qk::spa_two_atanh_f64(1.0), inf, "atanh(1)"}, {qk::spa_two_atanh_f64(-1.0), -inf, "atanh(-1)"},
        {qk::spa_two_atanh_f64(0.0), 0.0, "atanh(+0)"}, {qk::spa_two_atanh_f64(-0.0), -0.0, "atanh(-0)"},
        {qk::spa_two_atanh_f64(nan), nan, "atanh(nan)"}, {qk::spa_two_atanh_f64(1.5), nan, "atanh(1.5)"},
    };
    for (auto &s : sp) {
        const bool ok = (std::isnan(s.want) && std::isnan(s.got)) || (s.got == s.want && std::signbit(s.got) == std::signbit(s.want));
        printf("%-16s got %.17g want %.17g %s\n", s.what, s.got, s.want, ok ? "ok" : "MISMATCH");
        bad += !ok;
    }
    // where does tanh(x/2) become exactly 1.0?
    for (double x = 36.0; x < 39.0; x += 0.25) {
        const double a = qk::spa_tanh_half_f64(x), b = std::tanh(x / 2);
        printf("x=%.2f ours 1-%.3g libm 1-%.3g\n", x, 1 - a, 1 - b);
    }
    printf(bad ? "FAILED (%d)\n" : "all within 5.5 ulp\n", bad);
    return bad != 0;
}
