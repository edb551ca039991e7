/*
 * distnet_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's distributional value network
 * (model/model_distributional.py:18-52 Net): conv 4x4 (1->32) -> LeakyReLU(0.01) -> conv 4x4 (32->32) -> LeakyReLU ->
 * flatten -> fc 128 -> LeakyReLU -> fc `atoms` -> softmax.  The reference hard-codes a 22x10 input
 * (model_distributional.py:27: convOutShape((22, 10), ...)), so the 20x10 observation is given two empty rows on top
 * (the hidden rows of the 22-row era); flatten = 32 x 16 x 4 = 2048.
 * weights (state_dict order, PyTorch layouts): seq.conv1.w[32,1,4,4] b[32] seq.conv2.w[32,32,4,4] b[32]
 *   seq.fc1.w[128,2048] b[128] seq.fc_v.w[atoms,128] b[atoms].
 * Double accumulation, float layer outputs; pinned against the reference module through tests/golden/distnet_golden.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static float leaky(float x) { return x > 0 ? x : 0.01f * x; }

static void conv4(const float *in, int cin, int h, int w, const float *wt, const float *b, float *out) {
    int oh = h - 3, ow = w - 3;
    for (int co = 0; co < 32; ++co)
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                double acc = b[co];
                for (int ci = 0; ci < cin; ++ci)
                    for (int dy = 0; dy < 4; ++dy)
                        for (int dx = 0; dx < 4; ++dx) acc += (double)in[(ci * h + y + dy) * w + x + dx] * (double)wt[((co * cin + ci) * 4 + dy) * 4 + dx];
                out[(co * oh + y) * ow + x] = leaky((float)acc);
            }
}

void dn_forward(const float *W, const int8_t *states, int k, int bins, float *dist) {
    const float *c1w = W, *c1b = c1w + 512, *c2w = c1b + 32, *c2b = c2w + 16384, *f1w = c2b + 32, *f1b = f1w + 128 * 2048;
    const float *fvw = f1b + 128, *fvb = fvw + (size_t)bins * 128;
    float *a0 = (float *)calloc(220, 4), *a1 = (float *)malloc(4 * 32 * 19 * 7), *a2 = (float *)malloc(4 * 2048);
    float h[128], logit[256];
    for (int i = 0; i < k; ++i) {
        for (int j = 0; j < 200; ++j) a0[20 + j] = (float)states[i * 200 + j];      /* rows 0-1 stay empty */
        conv4(a0, 1, 22, 10, c1w, c1b, a1);
        conv4(a1, 32, 19, 7, c2w, c2b, a2);
        for (int o = 0; o < 128; ++o) {
            double acc = f1b[o];
            for (int j = 0; j < 2048; ++j) acc += (double)a2[j] * (double)f1w[(size_t)o * 2048 + j];
            h[o] = leaky((float)acc);
        }
        float mx = -INFINITY;
        for (int o = 0; o < bins; ++o) {
            double acc = fvb[o];
            for (int j = 0; j < 128; ++j) acc += (double)h[j] * (double)fvw[(size_t)o * 128 + j];
            logit[o] = (float)acc;
            if (logit[o] > mx) mx = logit[o];
        }
        double sum = 0;
        for (int o = 0; o < bins; ++o) sum += exp((double)logit[o] - (double)mx);
        for (int o = 0; o < bins; ++o) dist[(size_t)i * bins + o] = (float)(exp((double)logit[o] - (double)mx) / sum);
    }
    free(a0); free(a1); free(a2);
}
