// mlp64.cuh — model/mlp on the device (BASELINE configs[0]): the float64 scikit-learn-style
// MLP classifier the reference's default path trains (main.go:42-52 → model/mlp/mlp.go:45-65 →
// nn/neural_network/basemlp64.go `BaseMultilayerPerceptron64`: fit :484, fitStochastic :729, backprop :340,
// forwardPass :259, AdamOptimizer64.updateParams :1075).
//
// One hidden layer or more, biases, relu / logistic hidden activation, logistic output + binary_log_loss with the
// probabilities clipped to (nextafter(0,1), nextafter(1,0)) (:180-195), L2 term alpha/(2n)·Σw² in the loss (:361) and
// alpha/n·w in the coefficient gradients (:327), Adam whose beta powers advance once per parameter ELEMENT
// (:1082-1090 — the bias correction is gone within the first call), minibatches of 200 visited through a per-epoch
// shuffle (:788), tol / n_iter_no_change stopping (:826-840,886-890).
//
// All arithmetic is float64 (the reference's: gonum Dgemm).  A step is ~17 MFLOP — launch-latency, not throughput,
// bound: four kernels per minibatch, each thread owns one output element and walks its reduction in the oracle's
// order (k ascending), so the only differences against the CPU restatement are FMA contraction and libm ulps.
// Parameter layout = the reference's packed vector (:459-463): per layer [intercepts(fo) | coefs(fi x fo row-major)].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctr {

constexpr int kMlpMaxLayers = 8;
enum { MLP_ACT_RELU = 0, MLP_ACT_LOGISTIC = 1, MLP_ACT_IDENTITY = 2 };

struct Mlp64Dims {
    int n_layers;                    // len(layerUnits) = hidden layers + 2
    int units[kMlpMaxLayers];        // [nFeatures, hidden..., nOutputs]
    long off[kMlpMaxLayers];         // offset of layer l's [intercepts | coefs] in the packed vector
    int hidden_act;
};

// hidden / output layer l -> l+1 for the m rows of the minibatch: out[r, j] = act(b[j] + sum_k in[r, k] * W[k, j]).
// in == nullptr: layer 0 reads the dataset rows X[idx[r]] (the shuffled visiting order, basemlp64.go:788-808).
__global__ void __launch_bounds__(256)
k_mlp64_layer(const double* __restrict__ params, Mlp64Dims d, int l, const double* __restrict__ X, long ldx, const long* __restrict__ idx,
              const double* __restrict__ in, double* __restrict__ out, int m) {
    const int fi = d.units[l], fo = d.units[l + 1];
    const double* b = params + d.off[l];
    const double* W = b + fo;
    const bool last = (l + 1 == d.n_layers - 1);
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < (long)m * fo; t += (long)gridDim.x * blockDim.x) {
        const int r = (int)(t / fo), j = (int)(t % fo);
        const double* a = in ? in + (long)r * fi : X + (idx ? idx[r] : (long)r) * ldx;
        double z = 0.0;
        for (int k = 0; k < fi; k++) z += a[k] * W[(long)k * fo + j];
        z += b[j];                                                              // addIntercepts64 :205
        if (last || d.hidden_act == MLP_ACT_LOGISTIC) z = 1.0 / (1.0 + exp(-z));   // :82-88, out_activation :423-425
        else if (d.hidden_act == MLP_ACT_RELU) z = z < 0.0 ? 0.0 : z;             // :96-104
        out[t] = z;
    }
}

// loss of the minibatch (binary_log_loss :180-195 + L2 :361), delta of the output layer (h - y, :373-381), and the
// epoch accumulator acc += loss * m (:805).  One block; fixed-order reductions.
__global__ void __launch_bounds__(256)
k_mlp64_loss(const double* __restrict__ params, Mlp64Dims d, const double* __restrict__ h, const float* __restrict__ Y, const long* __restrict__ idx,
             int m, double alpha, double* __restrict__ delta, double* __restrict__ loss_out, double* __restrict__ epoch_acc) {
    __shared__ double red[256];
    const int no = d.units[d.n_layers - 1];
    const double hmin = 4.9406564584124654e-324, hmax = 0.99999999999999989;     // nextafter(0,1), nextafter(1,0)
    double s = 0.0;
    for (int i = threadIdx.x; i < m * no; i += 256) {
        const int r = i / no, j = i % no;
        const double y = (double)Y[(idx ? idx[r] : (long)r) * no + j];
        double hv = h[i];
        delta[i] = hv - y;
        hv = hv < hmin ? hmin : (hv > hmax ? hmax : hv);
        s += -y * log(hv) - (1.0 - y) * log1p(-hv);
    }
    red[threadIdx.x] = s; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    const double ll = red[0] / (double)m;
    __syncthreads();
    double sq = 0.0;                                                              // sumCoefSquares :311-319
    for (int l = 0; l + 1 < d.n_layers; l++) {
        const double* W = params + d.off[l] + d.units[l + 1];
        const long cnt = (long)d.units[l] * d.units[l + 1];
        for (long i = threadIdx.x; i < cnt; i += 256) sq += W[i] * W[i];
    }
    red[threadIdx.x] = sq; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        const double loss = ll + (0.5 * alpha) * red[0] / (double)m;
        *loss_out = loss;
        if (epoch_acc) *epoch_acc += loss * (double)m;
    }
}

// gradients of layer l (computeLossGrad :322-331): coefGrads = actᵀ·delta / m + alpha/m · coefs, interceptGrads =
// column means of delta; thread (k, j) walks the m rows in order.  k == fi addresses the intercept row.
__global__ void __launch_bounds__(256)
k_mlp64_grads(const double* __restrict__ params, double* __restrict__ grads, Mlp64Dims d, int l, const double* __restrict__ X, long ldx,
              const long* __restrict__ idx, const double* __restrict__ act, const double* __restrict__ delta, int m, double alpha) {
    const int fi = d.units[l], fo = d.units[l + 1];
    const double* W = params + d.off[l] + fo;
    double* gb = grads + d.off[l];
    double* gW = gb + fo;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < (long)(fi + 1) * fo; t += (long)gridDim.x * blockDim.x) {
        const int k = (int)(t / fo), j = (int)(t % fo);
        double s = 0.0;
        if (k == fi) {
            for (int r = 0; r < m; r++) s += delta[(long)r * fo + j];                  // matRowMean64 :213-226
            gb[j] = s / (double)m;
        } else {
            for (int r = 0; r < m; r++) {
                const double a = act ? act[(long)r * fi + k] : X[(idx ? idx[r] : (long)r) * ldx + k];
                s += a * delta[(long)r * fo + j];
            }
            gW[(long)k * fo + j] = s / (double)m + alpha / (double)m * W[(long)k * fo + j];
        }
    }
}

// delta of hidden layer l from delta of layer l+1 (:388-396): (delta_{l+1} · Wᵀ) ⊙ act'(z)
__global__ void __launch_bounds__(256)
k_mlp64_delta(const double* __restrict__ params, Mlp64Dims d, int l, const double* __restrict__ act, const double* __restrict__ dnext,
              double* __restrict__ dcur, int m) {
    const int fi = d.units[l], fo = d.units[l + 1];
    const double* W = params + d.off[l] + fo;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < (long)m * fi; t += (long)gridDim.x * blockDim.x) {
        const int r = (int)(t / fi), k = (int)(t % fi);
        double s = 0.0;
        for (int j = 0; j < fo; j++) s += dnext[(long)r * fo + j] * W[(long)k * fo + j];
        const double z = act[t];
        if (d.hidden_act == MLP_ACT_RELU) { if (z == 0.0) s = 0.0; }                   // :140-148
        else if (d.hidden_act == MLP_ACT_LOGISTIC) s *= z * (1.0 - z);                 // :124-131
        dcur[t] = s;
    }
}

// AdamOptimizer64.updateParams (:1075-1091).  The reference multiplies beta1t / beta2t once per ELEMENT, in packed
// order, across calls: element i of call t (1-based) sees beta^((t-1)*np + i + 1).  Evaluated here in closed form.
__global__ void __launch_bounds__(256)
k_mlp64_adam(double* __restrict__ params, const double* __restrict__ grads, double* __restrict__ ms, double* __restrict__ vs, long np,
             double calls_before, double lr_init, double b1, double b2, double eps) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < np; i += (long)gridDim.x * blockDim.x) {
        const double g = grads[i];
        const double m = b1 * ms[i] + (1.0 - b1) * g;
        const double v = b2 * vs[i] + (1.0 - b2) * g * g;
        ms[i] = m; vs[i] = v;
        const double e = calls_before * (double)np + (double)(i + 1);
        const double b1t = pow(b1, e), b2t = pow(b2, e);
        const double lr = lr_init * sqrt(1.0 - b2t) / (1.0 - b1t);
        params[i] += -lr * m / (sqrt(v) + eps);
    }
}

__global__ void __launch_bounds__(256)
k_f32_to_f64(const float* __restrict__ a, double* __restrict__ b, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = (double)a[i];
}
__global__ void __launch_bounds__(256)
k_f64_to_f32(const double* __restrict__ a, float* __restrict__ b, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = (float)a[i];
}

}  // namespace ctr
