// ccgp_math.h -- small FP64 math helpers shared by every kernel (host+device).
//
// dexp_neg(s) = exp(-s) for s >= 0: the Gaussian correlation exp(-Dist) of
// corr.matrix ([A]:351-360 `return(exp(-Dist))`).  Cody-Waite reduction by ln2
// in two pieces, then exp(r) = 1 + r + r^2 q(r) with a degree-9 near-minimax q
// on |r| <= ln2/2 (approximation error 1.6e-17, fitted in 60-digit mpmath at
// Chebyshev nodes), scaled by 2^k through the exponent field.  No special-case
// paths: the argument is always finite and <= 0; results below 2^-1000 flush
// to zero (irrelevant next to the unit diagonal).  14 FMA-pipe ops + 3 integer.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define CCGP_HD __host__ __device__ __forceinline__
#else
#define CCGP_HD static inline
#endif

CCGP_HD double ccgp_scale2(double v, int k) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    u += ((uint64_t)(int64_t)k) << 52;
    memcpy(&v, &u, 8);
    return v;
#endif
}

#if defined(__CUDACC__)
// Coefficients live in constant memory so each Horner step is one DFMA with a constant-bank
// operand (64-bit immediates would cost two extra uniform moves per step).
static __constant__ double c_dexp[16] = {
    6755399441055744.0, 1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    2.5100424157005067e-08, 2.7620138719733994e-07, 2.7557268378684192e-06, 2.480152119021773e-05,
    0.00019841269863105968, 0.0013888888917281794, 0.008333333333330051, 0.04166666666662399,
    0.16666666666666669, 0.5000000000000001, 700.0, 0.0};

// device version, branch-free.  CLAMP=false: the caller guarantees s < 1e8 (checked once per
// candidate from the parameter row and the design's bounding box), so only the power-of-two
// exponent is clamped (one integer max); CLAMP=true clamps the argument itself at 700.
template <bool CLAMP>
__device__ __forceinline__ double dexp_neg_dev(double s) {
    if (CLAMP) s = fmin(s, c_dexp[14]);
    double t = fma(-s, c_dexp[1], c_dexp[0]);
    double kf = t - c_dexp[0];
    double r = fma(kf, c_dexp[2], -s);
    r = fma(kf, c_dexp[3], r);
    double q = c_dexp[4];
    q = fma(q, r, c_dexp[5]);
    q = fma(q, r, c_dexp[6]);
    q = fma(q, r, c_dexp[7]);
    q = fma(q, r, c_dexp[8]);
    q = fma(q, r, c_dexp[9]);
    q = fma(q, r, c_dexp[10]);
    q = fma(q, r, c_dexp[11]);
    q = fma(q, r, c_dexp[12]);
    q = fma(q, r, c_dexp[13]);
    double r2 = r * r;
    double e = fma(r2, q, r) + 1.0;
    // t = MAGIC + k exactly, so k sits in the low word of t (two's complement)
    int k = __double2loint(t);
    if (!CLAMP) k = max(k, -1010);
    return __hiloint2double(__double2hiint(e) + (k << 20), __double2loint(e));
}
#endif

// ---- table-driven variant (the matrix-build loops) ---------------------------------------------
// exp(-s) = 2^e * T[j] * exp(r),  -s = (128 e + j) ln2/128 + r,  |r| <= ln2/256: the 128-entry table
// T[j] = 2^(j/128) (correctly rounded, 50-digit mpmath) shortens the polynomial to degree 5
// (truncation r^6/720 < 6e-19): 10 FP64 operations instead of 17, plus one 8-byte table load.
// Valid for 0 <= s < 1e6 (the integer k = round(-s 128/ln2) must fit an int32); the caller clamps.
#define CCGP_EXPT_INVL 184.6649652337873          /* 128 / ln2 */
#define CCGP_EXPT_LHI (-0.005415212348452769)     /* -(ln2/128), 32 significant bits: k * LHI is exact */
#define CCGP_EXPT_LLO (3.2819649005320973e-13)    /* -(ln2/128 - hi) */
#if defined(__CUDA_ARCH__)
#define CCGP_TABCONST __device__ const
#else
#define CCGP_TABCONST static const
#endif
CCGP_TABCONST double CCGP_EXP2_TAB[128] = {
    1.0, 1.0054299011128027, 1.0108892860517005, 1.016378314910953,
    1.0218971486541166, 1.0274459491187637, 1.0330248790212284, 1.0386341019613787,
    1.0442737824274138, 1.0499440858006872, 1.0556451783605572, 1.061377227289262,
    1.0671404006768237, 1.0729348675259756, 1.0787607977571199, 1.0846183622133092,
    1.0905077326652577, 1.0964290818163769, 1.102382583307841, 1.1083684117236787,
    1.1143867425958924, 1.1204377524096067, 1.1265216186082418, 1.1326385195987192,
    1.1387886347566916, 1.1449721444318042, 1.1511892299529827, 1.1574400736337511,
    1.1637248587775775, 1.1700437696832502, 1.1763969916502812, 1.182784710984341,
    1.189207115002721, 1.1956643920398273, 1.202156731452703, 1.2086843236265816,
    1.215247359980469, 1.2218460329727576, 1.22848053610687, 1.2351510639369334,
    1.241857812073484, 1.2486009771892048, 1.255380757024691, 1.2621973503942507,
    1.2690509571917332, 1.275941778396392, 1.2828700160787783, 1.2898358734066657,
    1.2968395546510096, 1.3038812651919358, 1.3109612115247644, 1.318079601266064,
    1.3252366431597413, 1.3324325470831615, 1.339667524053303, 1.3469417862329458,
    1.3542555469368927, 1.3616090206382248, 1.3690024229745905, 1.3764359707545302,
    1.383909881963832, 1.3914243757719262, 1.3989796725383112, 1.4065759938190154,
    1.4142135623730951, 1.4218926021691656, 1.42961333839197, 1.4373759974489824,
    1.4451808069770467, 1.4530279958490526, 1.460917794180647, 1.4688504333369818,
    1.4768261459394993, 1.4848451658727524, 1.4929077282912648, 1.5010140696264256,
    1.5091644275934228, 1.5173590411982147, 1.5255981507445384, 1.533881997840956,
    1.5422108254079407, 1.550584877685, 1.559004400237837, 1.567469639965553,
    1.5759808451078865, 1.5845382652524937, 1.593142151342267, 1.6017927556826934,
    1.6104903319492543, 1.6192351351948637, 1.6280274218573478, 1.6368674497669644,
    1.645755478153965, 1.6546917676561943, 1.6636765803267364, 1.6727101796415966,
    1.681792830507429, 1.6909247992693053, 1.7001063537185235, 1.709337763100463,
    1.718619298122478, 1.7279512309618377, 1.7373338352737062, 1.746767386199169,
    1.7562521603732995, 1.7657884359332727, 1.7753764925265212, 1.785016611318935,
    1.7947090750031072, 1.804454167806624, 1.8142521755003989, 1.8241033854070534,
    1.8340080864093424, 1.843966568958626, 1.8539791250833855, 1.864046048397789,
    1.8741676341103, 1.8843441790323345, 1.8945759815869656, 1.9048633418176741,
    1.9152065613971474, 1.925605943636125, 1.9360617934922943, 1.9465744175792332,
    1.9571441241754002, 1.9677712232331759, 1.978456026387951, 1.9891988469672663,
};
CCGP_HD double dexp_neg_tab(double s, const double* T) {
    const double MAGIC = 6755399441055744.0;
    if (s > 700.0) s = 700.0;
    double t = fma(-s, CCGP_EXPT_INVL, MAGIC);
    double kf = t - MAGIC;
    double r = fma(kf, CCGP_EXPT_LHI, -s);
    r = fma(kf, CCGP_EXPT_LLO, r);
    const int k = (int)kf;
    const double Tv = T[k & 127];
    const int e = k >> 7;
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    return ccgp_scale2(fma(Tv, p, Tv), e);
}
#if defined(__CUDACC__)
static __constant__ double c_dexpt[8] = {6755399441055744.0, CCGP_EXPT_INVL, CCGP_EXPT_LHI, CCGP_EXPT_LLO,
                                         1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 700.0};
// device version; T points to a shared-memory copy of CCGP_EXP2_TAB.  CLAMP=false: caller guarantees s < 1e6.
template <bool CLAMP>
__device__ __forceinline__ double dexp_neg_tab_dev(double s, const double* T) {
    if (CLAMP) s = fmin(s, c_dexpt[7]);
    const double t = fma(-s, c_dexpt[1], c_dexpt[0]);
    const double kf = t - c_dexpt[0];
    double r = fma(kf, c_dexpt[2], -s);
    r = fma(kf, c_dexpt[3], r);
    const int k = __double2loint(t);
    const double Tv = T[k & 127];
    int e = k >> 7;
    if (!CLAMP) e = max(e, -1010);
    double q = fma(r, c_dexpt[4], c_dexpt[5]);
    q = fma(q, r, c_dexpt[6]);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    const double v = fma(Tv, p, Tv);
    return __hiloint2double(__double2hiint(v) + (e << 20), __double2loint(v));
}
// the same with the clamp s <= ~700 done on the high word (one integer min instead of DSETP + two selects): for
// s >= 0 the order of doubles is the order of their high words; identical bits for s < 700, exp(-700.0005) ~ 1e-304
// at most beyond; negative and NaN arguments behave as with fmin (negative kept, NaN -> the clamp value).
// s <= 700.0005 keeps k >= -129267, so the exponent step k >> 7 >= -1010 needs no guard of its own.
// STRIDE: the table is replicated STRIDE times (entry j of copy c at T0[j * STRIDE + c]); the caller passes T = T0 + its
// copy, so the lanes of a half-warp hit different banks whatever their arguments (one table, random j: ~5 extra wavefronts)
template <int STRIDE>
__device__ __forceinline__ double dexp_neg_tab_dev_ic(double s, const double* T) {
    s = __hiloint2double(min(__double2hiint(s), 0x4085E000), __double2loint(s));
    const double t = fma(-s, c_dexpt[1], c_dexpt[0]);
    const double kf = t - c_dexpt[0];
    double r = fma(kf, c_dexpt[2], -s);
    r = fma(kf, c_dexpt[3], r);
    const int k = __double2loint(t);
    const double Tv = T[(k & 127) * STRIDE];
    double q = fma(r, c_dexpt[4], c_dexpt[5]);
    q = fma(q, r, c_dexpt[6]);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    const double v = fma(Tv, p, Tv);
    return __hiloint2double(__double2hiint(v) + ((k >> 7) << 20), __double2loint(v));
}
#endif

CCGP_HD double dexp_neg(double s) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() by add/sub
    double t = fma(-s, 1.4426950408889634074, MAGIC);
    double kf = t - MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, -s);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double q = 2.5100424157005067e-08;
    q = fma(q, r, 2.7620138719733994e-07);
    q = fma(q, r, 2.7557268378684192e-06);
    q = fma(q, r, 2.480152119021773e-05);
    q = fma(q, r, 0.00019841269863105968);
    q = fma(q, r, 0.0013888888917281794);
    q = fma(q, r, 0.008333333333330051);
    q = fma(q, r, 0.04166666666662399);
    q = fma(q, r, 0.16666666666666669);
    q = fma(q, r, 0.5000000000000001);
    double r2 = r * r;
    double e = fma(r2, q, r) + 1.0;
    int k = (int)kf;
    if (s > 690.0) return 0.0;
    return ccgp_scale2(e, k);
}


// ------------------------------------------------------------------------------------------
// 1-D component kernels of the reference's one-dimensional scripts:
//   Matern.corr.func  ([D1]:348-351 = "1D Combined GP Public.R"): t^nu K_nu(t) / (Gamma(nu) 2^(nu-1)),
//                     t = 2 sqrt(nu) |h| / theta, 1 at h = 0
//   spline.corr.func  ([D2]:346-357 = "1D Combined GP Two Families Public.R"): cubic spline, support theta
// K_nu for integer nu comes from K0, K1 and the upward recurrence K_{m+1} = K_{m-1} + (2m/t) K_m
// (all terms positive: stable); half-integer nu starts from K_{1/2} = sqrt(pi/2t) e^-t.
// K0/K1: exact power series for t <= 2, Chebyshev expansion of sqrt(2t/pi) e^t K(t) in 4/t - 1
// beyond (tables from tools/gen_bessel_coeffs.py, 50-digit mpmath).
#if defined(__CUDA_ARCH__)
#define CCGP_CONST __device__ const
#else
#define CCGP_CONST static const
#endif
// I0 series: sum a0[k] z^k, z = t^2/4
CCGP_CONST double BES_I0[16] = {1.0, 1.0, 0.25, 0.027777777777777776, 0.001736111111111111, 6.944444444444444e-05, 1.9290123456790124e-06, 3.936759889140842e-08, 6.151187326782565e-10, 7.594058428126624e-12, 7.594058428126623e-14, 6.276081345559193e-16, 4.358389823304995e-18, 2.5789288895295828e-20, 1.3157800456783586e-22, 5.8479113141260385e-25};
CCGP_CONST double BES_K0P[16] = {-0.5772156649015329, 0.42278433509846713, 0.23069608377461678, 0.0348921574564389, 0.0026147876188052093, 0.00011848039364109726, 3.6126241031992037e-06, 7.935096521304209e-08, 1.3167486730385647e-09, 1.709994072705808e-11, 1.785934656987074e-13, 1.5330343403208473e-15, 1.1009270959725744e-17, 6.712740659979047e-20, 3.518851972639805e-22, 1.6029202854896424e-24};
CCGP_CONST double BES_I1[16] = {1.0, 0.5, 0.08333333333333333, 0.006944444444444444, 0.00034722222222222224, 1.1574074074074073e-05, 2.755731922398589e-07, 4.920949861426052e-09, 6.834652585313961e-11, 7.594058428126623e-13, 6.903689480115112e-15, 5.230067787965994e-17, 3.352607556388458e-19, 1.842092063949702e-21, 8.771866971189057e-24, 3.654944571328774e-26};
CCGP_CONST double BES_K1P[16] = {-0.15443132980306573, 0.6727843350984671, 0.18157516696085563, 0.019182189839330562, 0.001115359491966528, 4.142247689271143e-05, 1.071545914091181e-06, 2.045286003593878e-08, 3.002048746589188e-10, 3.495928729692882e-12, 3.309914735250272e-14, 2.5986411321011286e-16, 1.7195232826992565e-18, 9.721207518823618e-21, 4.750281743327667e-23, 2.0264937604328578e-25};
CCGP_CONST double BES_K0A[26] = {1.9470801528600736, -0.02509195450338081, 0.0012525861146772193, -0.00010252457224451742, 1.1130340992367562e-05, -1.4615294507429678e-06, 2.2075978855319634e-07, -3.718532935142939e-08, 6.841089366293377e-09, -1.3544365764715988e-09, 2.8543500586874657e-10, -6.349157810923372e-11, 1.480833144458278e-11, -3.602127949377824e-12, 9.09860149387498e-13, -2.3777733246760547e-13, 6.409319528042818e-14, -1.7772984923934982e-14, 5.05864913473057e-15, -1.4749641154455083e-15, 4.3979843802055255e-16, -1.3390347046986472e-16, 4.157291155115988e-17, -1.3145791186184579e-17, 4.229134271580504e-18, -1.3828705421726237e-18};  // Chebyshev, c[0] is halved at evaluation
CCGP_CONST double BES_K1A[26] = {2.170745633103452, 0.0829191449155865, -0.0022802079498951454, 0.00015575944821741804, -1.5448624702449026e-05, 1.9200971856838045e-06, -2.7941602977436566e-07, 4.580722385967015e-08, -8.254724141098338e-09, 1.6077770889213073e-09, -3.343419366765729e-10, 7.35515136480485e-11, -1.6994684532881865e-11, 4.100839143561442e-12, -1.0286119996309399e-12, 2.671652354631771e-13, -7.162374471604971e-14, 1.9764832698093292e-14, -5.6010196328357906e-15, 1.6266497804027024e-15, -4.832824501299199e-16, 1.4665864849973659e-16, -4.539534566633093e-17, 1.4314456324007267e-17, -4.593217542733376e-18, 1.4983196424996574e-18};  // Chebyshev, c[0] is halved at evaluation

CCGP_HD void ccgp_bessel_k01(double t, double* k0, double* k1) {
    if (t <= 2.0) {
        const double z = 0.25 * t * t;
        double i0 = 0.0, p0 = 0.0, i1 = 0.0, p1 = 0.0;
        for (int k = 15; k >= 0; --k) {
            i0 = fma(i0, z, BES_I0[k]);
            p0 = fma(p0, z, BES_K0P[k]);
            i1 = fma(i1, z, BES_I1[k]);
            p1 = fma(p1, z, BES_K1P[k]);
        }
        const double lg = log(0.5 * t);
        *k0 = fma(-lg, i0, p0);
        *k1 = 1.0 / t + 0.5 * t * (lg * i1 - 0.5 * p1);
    } else {
        const double s = 4.0 / t - 1.0, s2 = 2.0 * s;
        double b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0;
        for (int k = 25; k >= 1; --k) {
            const double tb = fma(s2, b0, BES_K0A[k]) - b1; b1 = b0; b0 = tb;
            const double tc = fma(s2, c0, BES_K1A[k]) - c1; c1 = c0; c0 = tc;
        }
        const double f0 = fma(s, b0, 0.5 * BES_K0A[0]) - b1;
        const double f1 = fma(s, c0, 0.5 * BES_K1A[0]) - c1;
        const double sc = sqrt(1.5707963267948966 / t) * exp(-t);
        *k0 = sc * f0;
        *k1 = sc * f1;
    }
}

// Matern correlation at scaled distance t = 2 sqrt(nu) |h| / theta; twonu = 2 nu (integer),
// norm = 1 / (Gamma(nu) 2^(nu-1))
CCGP_HD double ccgp_matern(double t, int twonu, double norm) {
    if (!(t > 0.0)) return 1.0;
    if (t > 700.0) return 0.0;
    double km, kc, mu;          // K_{mu-1}, K_mu
    if (twonu & 1) {
        const double sc = sqrt(1.5707963267948966 / t) * exp(-t);
        km = sc; kc = sc * (1.0 + 1.0 / t); mu = 1.5;          // K_{1/2}, K_{3/2}
        if (twonu == 1) kc = km;
    } else {
        ccgp_bessel_k01(t, &km, &kc); mu = 1.0;                 // K_0, K_1
        if (twonu == 0) kc = km;
    }
    const double two_over_t = 2.0 / t;
    for (int m = (twonu & 1) ? 3 : 2; m < twonu; m += 2) {      // raise the order to nu
        const double kn = fma(mu * two_over_t, kc, km);
        km = kc; kc = kn; mu += 1.0;
    }
    double pw = 1.0;
    for (int m = 0; m < (twonu >> 1); ++m) pw *= t;
    if (twonu & 1) pw *= sqrt(t);
    return pw * kc * norm;
}

// cubic-spline correlation at u = |h| / theta
CCGP_HD double ccgp_spline(double u) {
    if (u <= 0.5) return fma(6.0 * u * u, u - 1.0, 1.0);
    if (u <= 1.0) { const double v = 1.0 - u; return 2.0 * v * v * v; }
    return 0.0;
}
