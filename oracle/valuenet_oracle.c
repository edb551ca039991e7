/*
 * valuenet_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference value network forward pass
 * (model/model_vv.py:13-52 Net, :210-217 Model_VV.inference): 3 x [conv3x3(32), valid, ReLU] 20x10 -> 18x8 ->
 * 16x6 -> 14x4, NCHW flatten (1792) -> fc 256 -> ReLU -> fc 2 -> sigmoid -> * out_ubound + out_lbound.
 * Pinned against the reference's own Model_VV (torch CPU) through tests/golden/valuenet_*.npz.
 * Accumulates in double and rounds each layer's output to float, so it sits within fp32 rounding of the
 * reference (which accumulates in fp32 in library-defined order); tolerance is stated in the tests (rtol 1e-5).
 *
 * weights: one flat float array, PyTorch layouts, in state_dict order:
 *   conv1.w[32,1,3,3] conv1.b[32] conv2.w[32,32,3,3] conv2.b[32] conv3.w[32,32,3,3] conv3.b[32]
 *   fc1.w[256,1792] fc1.b[256] fc_out.w[2,256] fc_out.b[2] out_ubound[2] out_lbound[2]      (478342 floats)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define VO_NW 478342

static void conv3x3(const float *in, int cin, int h, int w, const float *wt, const float *b, float *out) {
    int oh = h - 2, ow = w - 2;
    for (int co = 0; co < 32; ++co)
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                double acc = b[co];
                for (int ci = 0; ci < cin; ++ci)
                    for (int dy = 0; dy < 3; ++dy)
                        for (int dx = 0; dx < 3; ++dx)
                            acc += (double)in[(ci * h + y + dy) * w + x + dx] * (double)wt[((co * cin + ci) * 3 + dy) * 3 + dx];
                float r = (float)acc;
                out[(co * oh + y) * ow + x] = r > 0 ? r : 0;   /* ReLU, model_vv.py:22 */
            }
}

void vo_forward(const float *W, const int8_t *states, int k, float *v, float *var) {
    const float *c1w = W, *c1b = c1w + 288, *c2w = c1b + 32, *c2b = c2w + 9216, *c3w = c2b + 32, *c3b = c3w + 9216;
    const float *f1w = c3b + 32, *f1b = f1w + 458752, *fow = f1b + 256, *fob = fow + 512, *ub = fob + 2, *lb = ub + 2;
    float *a0 = (float *)malloc(sizeof(float) * 200), *a1 = (float *)malloc(sizeof(float) * 32 * 18 * 8);
    float *a2 = (float *)malloc(sizeof(float) * 32 * 16 * 6), *a3 = (float *)malloc(sizeof(float) * 1792);
    float h[256];
    for (int i = 0; i < k; ++i) {
        for (int j = 0; j < 200; ++j) a0[j] = (float)states[i * 200 + j];   /* model_vv.py:212 as_tensor(float) */
        conv3x3(a0, 1, 20, 10, c1w, c1b, a1);
        conv3x3(a1, 32, 18, 8, c2w, c2b, a2);
        conv3x3(a2, 32, 16, 6, c3w, c3b, a3);                               /* [32,14,4] == NCHW flatten order */
        for (int o = 0; o < 256; ++o) {
            double acc = f1b[o];
            for (int j = 0; j < 1792; ++j) acc += (double)a3[j] * (double)f1w[o * 1792 + j];
            float r = (float)acc;
            h[o] = r > 0 ? r : 0;
        }
        float out[2];
        for (int o = 0; o < 2; ++o) {
            double acc = fob[o];
            for (int j = 0; j < 256; ++j) acc += (double)h[j] * (double)fow[o * 256 + j];
            float pre = (float)acc;
            float s = (float)(1.0 / (1.0 + exp(-(double)pre)));            /* Sigmoid, model_vv.py:42 */
            out[o] = s * ub[o] + lb[o];                                     /* model_vv.py:51 */
        }
        v[i] = out[0]; var[i] = out[1];
    }
    free(a0); free(a1); free(a2); free(a3);
}
