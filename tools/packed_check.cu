// Is the packed (f32x2) node arithmetic bit-identical to the scalar fp32 arithmetic?  (development aid, GPU only)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/packed_check tools/packed_check.cu
#include <cstdio>
#include <cstring>
#include "../latticeboltzmannsimulations_b200/csrc/lbm_device.cuh"
using namespace lbm;

__device__ __forceinline__ unsigned long long mix(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    unsigned long long s;
    __device__ float uni() { s = mix(s); return (float)((s >> 40) * (1.0 / 16777216.0)); }          // [0,1)
    __device__ float sym() { return 2.0f * uni() - 1.0f; }
};
struct Bad { int combo; float in[2][11]; float a[13], b[13]; int lane; };

__device__ void make_node(Rng& r, float f[9], float& pi, float& rp) {
    const float scales[6] = {0.1f, 1e-2f, 1e-4f, 1e-6f, 3e-9f, 0.0f};
    const float su = scales[(int)(r.uni() * 6) % 6], sn = scales[(int)(r.uni() * 6) % 6];
    const float rho = 1.0f + 0.05f * r.sym(), ux = su * r.sym(), uy = su * r.sym();
    feq_all_ref<float>(rho, ux, uy, f);
    for (int k = 0; k < 9; ++k) f[k] += 0.01f * sn * r.sym();
    float fe[9];
    feq_all_ref<float>(rho, ux * (1 + 0.1f * r.sym()), uy, fe);
    pi = fe[5] - fe[6] + fe[7] - fe[8];
    rp = rho * (1.0f + 1e-3f * r.sym());
}

template <int COLL, bool NEED_U, bool TURB>
__device__ void one(int combo, const CavityParams& cp, const float fa[9], const float fb[9], float pia, float pib, float rpa,
                    float rpb, unsigned long long* nbad, Bad* bad) {
    const Rates<float> rt(cp);
    const Rates<f32x2> rta(cp);
    float A[9], B[9];
    f32x2 P[9];
    for (int k = 0; k < 9; ++k) { A[k] = fa[k]; B[k] = fb[k]; P[k] = f32x2(fa[k], fb[k]); }
    float ra, uxa, uya, rb, uxb, uyb, p1a = 0, p1b = 0, ia = 0, ib = 0;
    f32x2 rp_, uxp, uyp, p1p(0.0f), ip(0.0f);
    const float oma = TURB ? smagorinsky_omega<float>(A, pia, rpa, rt.tau0) : 0.0f;
    const float omb = TURB ? smagorinsky_omega<float>(B, pib, rpb, rt.tau0) : 0.0f;
    const f32x2 omp = TURB ? smagorinsky_omega<f32x2>(P, f32x2(pia, pib), f32x2(rpa, rpb), rta.tau0) : f32x2(0.0f);
    node_update<float, COLL, NEED_U, TURB>(A, rt, false, false, false, false, ra, uxa, uya, oma, &p1a, &ia);
    node_update<float, COLL, NEED_U, TURB>(B, rt, false, false, false, false, rb, uxb, uyb, omb, &p1b, &ib);
    node_update<f32x2, COLL, NEED_U, TURB>(P, rta, false, false, false, false, rp_, uxp, uyp, omp, &p1p, &ip);
    constexpr bool LEAN = COLL == COLL_MRT && !NEED_U && !TURB;
    float va[13], vb[13];
    for (int k = 0; k < 9; ++k) { va[k] = A[k]; vb[k] = P[k].v.x; }
    va[9] = LEAN ? 0 : ra; vb[9] = LEAN ? 0 : rp_.v.x; va[10] = LEAN ? 0 : uxa; vb[10] = LEAN ? 0 : uxp.v.x;
    va[11] = TURB ? p1a + ia : 0; vb[11] = TURB ? p1p.v.x + ip.v.x : 0; va[12] = oma; vb[12] = omp.v.x;
    float wa[13], wb[13];
    for (int k = 0; k < 9; ++k) { wa[k] = B[k]; wb[k] = P[k].v.y; }
    wa[9] = LEAN ? 0 : rb; wb[9] = LEAN ? 0 : rp_.v.y; wa[10] = LEAN ? 0 : uxb; wb[10] = LEAN ? 0 : uxp.v.y;
    wa[11] = TURB ? p1b + ib : 0; wb[11] = TURB ? p1p.v.y + ip.v.y : 0; wa[12] = omb; wb[12] = omp.v.y;
    bool ok0 = true, ok1 = true;
    for (int k = 0; k < 13; ++k) {
        ok0 = ok0 && __float_as_uint(va[k]) == __float_as_uint(vb[k]);
        ok1 = ok1 && __float_as_uint(wa[k]) == __float_as_uint(wb[k]);
    }
    if (!ok0 || !ok1) {
        const unsigned long long n = atomicAdd(&nbad[combo], 1ull);
        if (n == 0) {
            Bad& q = bad[combo];
            q.combo = combo; q.lane = ok0 ? 1 : 0;
            for (int k = 0; k < 9; ++k) { q.in[0][k] = fa[k]; q.in[1][k] = fb[k]; }
            q.in[0][9] = pia; q.in[0][10] = rpa; q.in[1][9] = pib; q.in[1][10] = rpb;
            for (int k = 0; k < 13; ++k) { q.a[k] = ok0 ? wa[k] : va[k]; q.b[k] = ok0 ? wb[k] : vb[k]; }
        }
    }
}

// intermediates of the closure's pi block, scalar vs packed (index of the first differing stage, -1 if none)
template <typename T> __device__ void pi_stages(T rho, T ux, T uy, T out[8]) {
    const T usqr = fm(ux, ux, uy * uy);
    const T r2 = rho * w_diag<T>();
    const T fe5 = feq_one(r2, ux + uy, usqr), fe6 = feq_one(r2, uy - ux, usqr);
    const T fe7 = feq_one(r2, -(ux + uy), usqr), fe8 = feq_one(r2, ux - uy, usqr);
    out[0] = usqr; out[1] = r2; out[2] = fe5; out[3] = fe6; out[4] = fe7; out[5] = fe8; out[6] = fe5 - fe6 + fe7;
    out[7] = fe5 - fe6 + fe7 - fe8;
}
__global__ void stages(unsigned long long seed, unsigned long long* hist) {
    Rng r{seed + (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 104729ull};
    for (int rep = 0; rep < 64; ++rep) {
        const float rho = 1.0f + 0.05f * r.sym(), ux = 0.1f * r.sym(), uy = 0.1f * r.sym();
        float a[8];
        f32x2 p[8];
        pi_stages<float>(rho, ux, uy, a);
        pi_stages<f32x2>(f32x2(rho, rho), f32x2(ux, ux), f32x2(uy, uy), p);
        int first = 8;
        for (int k = 7; k >= 0; --k) if (__float_as_uint(a[k]) != __float_as_uint(p[k].v.x)) first = k;
        atomicAdd(&hist[first], 1ull);
    }
}

__global__ void check(CavityParams cp, unsigned long long seed, unsigned long long* nbad, Bad* bad) {
    Rng r{seed + (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x) * 7919ull};
    for (int rep = 0; rep < 64; ++rep) {
        float fa[9], fb[9], pia, pib, rpa, rpb;
        make_node(r, fa, pia, rpa);
        make_node(r, fb, pib, rpb);
        one<0, false, false>(0, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<0, true, false>(1, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<0, false, true>(2, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<0, true, true>(3, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<1, false, false>(4, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<1, true, false>(5, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<1, false, true>(6, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<1, true, true>(7, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<2, false, false>(8, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<2, true, false>(9, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<2, false, true>(10, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
        one<2, true, true>(11, cp, fa, fb, pia, pib, rpa, rpb, nbad, bad);
    }
}

int main() {
    unsigned long long* nbad; Bad* bad;
    cudaMallocManaged(&nbad, 12 * sizeof(*nbad)); cudaMallocManaged(&bad, 12 * sizeof(Bad));
    memset(nbad, 0, 12 * sizeof(*nbad));
    const double visc[3] = {0.08 * 610 / 1000.0, 0.08 * 4096 / 5000.0, 0.1 * 64 / 100.0};     // tau = 0.65, 0.70, 0.69
    for (int i = 0; i < 3; ++i) {
        CavityParams cp;
        const double omega = 1.0 / (3 * visc[i] + 0.5);
        cp.uLB = 0.08; cp.omega = omega; cp.omegam = 1.0 / (0.25 / (1.0 / omega - 0.5) + 0.5);
        cp.s_e = 1.64; cp.s_eps = 1.54; cp.s_q = 1.9; cp.tau0 = 1.0 / omega; cp.pad = 0;
        check<<<148 * 8, 128>>>(cp, 12345 + 1000003ull * i, nbad, bad);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s, %lld node pairs per combo\n", cudaGetErrorString(e), 3 * 148LL * 8 * 128 * 64);
    const char* names[3] = {"SRT", "TRT", "MRT"};
    for (int c = 0; c < 12; ++c) {
        printf("%s need_u=%d turb=%d : %llu mismatching pairs\n", names[c / 4], (c >> 0) & 1, (c >> 1) & 1, nbad[c]);
        if (nbad[c]) {
            const Bad& q = bad[c];
            printf("   lane %d inputs:", q.lane);
            for (int k = 0; k < 11; ++k) printf(" %.9g", q.in[q.lane][k]);
            printf("\n   scalar:");
            for (int k = 0; k < 13; ++k) printf(" %.9g", q.a[k]);
            printf("\n   packed:");
            for (int k = 0; k < 13; ++k) printf(" %.9g", q.b[k]);
            printf("\n");
        }
    }
    unsigned long long* hist;
    cudaMallocManaged(&hist, 9 * sizeof(*hist));
    memset(hist, 0, 9 * sizeof(*hist));
    stages<<<148 * 8, 128>>>(777, hist);
    cudaDeviceSynchronize();
    const char* st[9] = {"usqr", "r2", "fe5", "fe6", "fe7", "fe8", "fe5-fe6+fe7", "pi", "none"};
    for (int k = 0; k < 9; ++k) printf("first differing stage %-12s : %llu\n", st[k], hist[k]);
    return 0;
}
