// Development harness ONLY: runs the lmdif restatement of csrc/lod.cu on the host CPU.
#include "../precise-mrd-mini_b200/csrc/lod.cu"
#include <cstdio>
int main(int argc, char** argv) {
    double af[15];
    for (int i = 0; i < 15; ++i) af[i] = pow(10.0, -4.0 + 2.0 * i / 14.0);
    double curves[3][15] = {{0,0,0,0,0,0,0,0,0,0,0,0,0,0,0},{1,1,1,1,1,1,1,1,1,1,1,1,1,1,1},{0,0,0,0,0,.04,0,.08,.2,.4,.6,.8,.92,1,1}};
    for (int c = 0; c < 3; ++c) {
        LmProblem P; P.m = 15;
        for (int i = 0; i < 15; ++i) { P.x[i] = log10(af[i]); P.y[i] = curves[c][i]; }
        double p[2] = {1.0, 1.0};
        int info = lm_lmdif(P, p, 2000);
        printf("curve %d: info=%d a=%.17g b=%.17g lod=%.17g\n", c, info, p[0], p[1], lod_fit_one(P, af, 0.95));
    }
    return 0;
}
