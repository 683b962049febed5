// dopt/nnet.cpp -- see nnet.hpp.
#include "nnet.hpp"

#include <cmath>
#include <cstdio>

namespace dopt {
namespace nnet {

static std::mt19937_64& rng() {
    static std::mt19937_64 g(1234);
    return g;
}
void seedInitializers(uint64_t seed) { rng().seed(seed); }

ParamInitializer constantInit(float v) {
    return [v](Operation param) {
        std::vector<float> vals(param->volume(), v);
        param->value()->set(vals.data(), vals.size() * sizeof(float));
    };
}
ParamInitializer heGaussianInit() {
    // parameters.d:262-272 via gaussianInit (parameters.d:53-68): stddev = sqrt(2 / fanIn)
    return [](Operation param) {
        size_t fanIn = 1;
        for (size_t i = 1; i < param->shape().size(); ++i) fanIn *= param->shape()[i];
        std::normal_distribution<float> dist(0.0f, std::sqrt(2.0f / (float)fanIn));
        std::vector<float> vals(param->volume());
        for (auto& x : vals) x = dist(rng());
        param->value()->set(vals.data(), vals.size() * sizeof(float));
    };
}

std::vector<LayerPtr> topologicalSort(const std::vector<LayerPtr>& layers) {
    std::vector<LayerPtr> sorted;
    std::map<const Layer*, bool> seen;
    std::function<void(const LayerPtr&)> visit = [&](const LayerPtr& l) {
        if (seen[l.get()]) return;
        seen[l.get()] = true;
        for (auto& d : l->deps()) visit(d);
        sorted.push_back(l);
    };
    for (auto& l : layers) visit(l);
    return sorted;
}

LayerPtr dataSource(Operation var) { return std::make_shared<Layer>(std::vector<LayerPtr>{}, var, var, std::vector<Parameter>{}); }
LayerPtr dataSource(Operation var, Operation trainVar) {
    return std::make_shared<Layer>(std::vector<LayerPtr>{}, var, trainVar, std::vector<Parameter>{});
}

static Operation safeAdd(Operation a, Operation b) {
    if (!a && !b) return nullptr;
    if (!a) return b;
    if (!b) return a;
    return a + b;
}

LayerPtr conv2D(LayerPtr input, size_t outputChannels, std::vector<size_t> filterDims, Conv2DOptions opts) {
    // nnet/layers/conv.d:74-165
    auto x = input->output();
    auto xTr = input->trainOutput();
    std::vector<size_t> fshape{outputChannels, x->shape()[1]};
    for (auto d : filterDims) fshape.push_back(d);
    auto filters = float32(fshape);
    opts.filterInit(filters);
    Operation filterLoss;
    filterLoss = safeAdd(filterLoss, opts.weightDecay == 0.0f ? nullptr : (opts.weightDecay * sum(filters * filters)));
    auto y = convolution(x, filters, opts.padding, opts.stride);
    auto yTr = (xTr == x) ? y : convolution(xTr, filters, opts.padding, opts.stride);
    std::vector<Parameter> params{Parameter{filters, filterLoss, opts.filterProj}};
    if (opts.useBias) {
        auto biases = float32(std::vector<size_t>{outputChannels});
        opts.biasInit(biases);
        auto yb = addBias(y, biases);
        yTr = (yTr == y) ? yb : addBias(yTr, biases);
        y = yb;
        params.push_back(Parameter{biases, nullptr, opts.biasProj});
    }
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, params);
}

LayerPtr dense(LayerPtr input, size_t numOutputs, DenseOptions opts) {
    // nnet/layers/dense.d:68-146
    auto x = input->output();
    auto xTr = input->trainOutput();
    bool same = (x == xTr);
    x = reshape(x, {x->shape()[0], x->volume() / x->shape()[0]});
    xTr = same ? x : reshape(xTr, {xTr->shape()[0], xTr->volume() / xTr->shape()[0]});
    auto weights = float32({numOutputs, x->shape()[1]});
    opts.weightInit(weights);
    Operation weightLoss;
    weightLoss = safeAdd(weightLoss, opts.weightDecay == 0.0f ? nullptr : (opts.weightDecay * sum(weights * weights)));
    auto y = matmul(x, transpose(weights, {1, 0}));
    auto yTr = same ? y : matmul(xTr, transpose(weights, {1, 0}));
    std::vector<Parameter> params{Parameter{weights, weightLoss, opts.weightProj}};
    if (opts.useBias) {
        auto bias = float32(std::vector<size_t>{numOutputs});
        opts.biasInit(bias);
        auto yb = y + repeat(bias, y->shape()[0]);
        yTr = same ? yb : (yTr + repeat(bias, yTr->shape()[0]));
        y = yb;
        params.push_back(Parameter{bias, nullptr, opts.biasProj});
    }
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, params);
}

LayerPtr batchNorm(LayerPtr input, BatchNormOptions opts) {
    // nnet/layers/batchnorm.d:68-156.  The running mean / variance come back packed behind the activations and are fed
    // to the optimiser as "projections" that overwrite `mean` and `var` (batchnorm.d:140-154).
    auto x = input->output();
    auto xTr = input->trainOutput();
    size_t C = x->shape()[1];
    auto gamma = float32({1, C, 1, 1});
    auto beta = float32(std::vector<size_t>{C});
    opts.gammaInit(gamma);
    opts.betaInit(beta);
    auto mean = float32(std::vector<size_t>{C});
    auto var = float32(std::vector<size_t>{C}, std::vector<float>(C, 1.0f));
    auto bnop = batchNormTrain(xTr, gamma, beta, mean, var, (double)opts.momentum);
    auto yTr = bnop[0];
    auto meanUpdateSym = bnop[1];
    auto varUpdateSym = bnop[2];
    auto y = batchNormInference(x, gamma, beta, mean, var);
    Projection meanUpdater = [meanUpdateSym](Operation) { return meanUpdateSym; };
    Projection varUpdater = [varUpdateSym](Operation) { return varUpdateSym; };
    std::vector<Parameter> params{
        Parameter{gamma, opts.gammaDecay == 0.0f ? nullptr : (opts.gammaDecay * sum(gamma * gamma)), opts.gammaProj},
        Parameter{beta, nullptr, opts.betaProj}, Parameter{mean, nullptr, meanUpdater}, Parameter{var, nullptr, varUpdater}};
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, params);
}

LayerPtr relu(LayerPtr input) {
    auto y = dopt::relu(input->output());
    auto yTr = input->output() == input->trainOutput() ? y : dopt::relu(input->trainOutput());
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, std::vector<Parameter>{});
}
LayerPtr maxPool(LayerPtr input, std::vector<size_t> dims) {
    auto y = dopt::maxpool(input->output(), dims);
    auto yTr = input->output() == input->trainOutput() ? y : dopt::maxpool(input->trainOutput(), dims);
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, std::vector<Parameter>{});
}
LayerPtr dropout(LayerPtr input, float dropProb) {
    // nnet/layers/dropout.d:14-29: train output = (uniform > dropProb) * x with a fresh mask per execution; test output =
    // x * (1 - dropProb).  dropMask and scale are plain variables filled with the constant, exactly as in the reference.
    auto x = input->output();
    auto xTr = input->trainOutput();
    auto dropMask = float32(xTr->shape(), std::vector<float>(xTr->volume(), dropProb));
    auto yTr = gt(uniformSample(xTr->shape()), dropMask) * xTr;
    auto scale = float32(x->shape(), std::vector<float>(x->volume(), 1.0f - dropProb));
    auto y = x * scale;
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, std::vector<Parameter>{});
}
LayerPtr softmax(LayerPtr input) {
    auto y = dopt::softmax(input->output());
    auto yTr = input->output() == input->trainOutput() ? y : dopt::softmax(input->trainOutput());
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, std::vector<Parameter>{});
}

DAGNetwork::DAGNetwork(std::vector<Operation> inputs, std::vector<LayerPtr> outputs) : mInputs(std::move(inputs)) {
    // nnet/networks.d:33-65
    for (auto& l : outputs) {
        mOutputs.push_back(l->output());
        mTrainOutputs.push_back(l->trainOutput());
    }
    for (auto& l : topologicalSort(outputs)) {
        for (auto& p : l->params()) {
            mParams.push_back(p.symbol);
            if (p.loss) mParameterLoss = mParameterLoss ? (mParameterLoss + p.loss) : p.loss;
            if (p.projection) mParameterProj[p.symbol] = p.projection;
        }
    }
    if (!mParameterLoss) mParameterLoss = float32({}, {0.0f});
}
void DAGNetwork::save(const std::string& filename) const {
    FILE* f = std::fopen(filename.c_str(), "wb");
    enforce(f != nullptr, "cannot open " + filename);
    for (auto& p : mParams) {
        auto v = p->value()->get<float>();
        std::fwrite(v.data(), sizeof(float), v.size(), f);
    }
    std::fclose(f);
}
void DAGNetwork::load(const std::string& filename) {
    FILE* f = std::fopen(filename.c_str(), "rb");
    enforce(f != nullptr, "cannot open " + filename);
    for (auto& p : mParams) {
        std::vector<float> v(p->volume());
        size_t got = std::fread(v.data(), sizeof(float), v.size(), f);
        if (got != v.size()) {
            std::fclose(f);
            throw Exception("parameter file is too short");
        }
        p->value()->set(v.data(), v.size() * sizeof(float));
    }
    std::fclose(f);
}

// ---- nnet/lipschitz.d ----------------------------------------------------------------------------------------------------------
Operation matrixNorm(Operation param, float p, size_t n) {
    enforce(param->rank() == 2, "This function only operates on matrices");
    if (p == 1.0f) {
        // maximum absolute ROW sum: dense / convolution weights are transposed before use (lipschitz.d:49-63)
        return maxElement(sum(abs(param), {1}));
    } else if (p == 2.0f) {
        // power iteration on W W^T from a random start (lipschitz.d:65-79)
        auto x = uniformSample({param->shape()[0], 1}) * 2.0f - 1.0f;
        auto weightsT = transpose(param, {1, 0});
        auto wwT = matmul(param, weightsT);
        for (size_t i = 0; i < n; ++i) x = matmul(wwT, x);
        auto v = x / sqrt(sum(x * x));
        auto y = matmul(weightsT, v);
        return sqrt(sum(y * y));
    } else if (std::isinf(p) && p > 0) {
        return maxElement(sum(abs(param), {0}));   // maximum absolute column sum (lipschitz.d:81-90)
    }
    throw Exception("Cannot compute matrix norm for p=" + std::to_string(p));
}

Operation convParamsNorm(Operation param, std::vector<size_t> inShape, std::vector<size_t> stride,
                         std::vector<size_t> padding, float p, size_t n) {
    if (p == 2.0f) {
        // power iteration on conv^T conv over a random image (lipschitz.d:114-128)
        std::vector<size_t> xs{1, param->shape()[1]};
        xs.insert(xs.end(), inShape.begin(), inShape.end());
        auto x = uniformSample(xs) * 2.0f - 1.0f;
        for (size_t i = 0; i < n; ++i) x = convolutionTranspose(convolution(x, param, padding, stride), param, padding, stride);
        auto v = x / sqrt(sum(x * x));
        auto y = convolution(v, param, padding, stride);
        return sqrt(sum(y * y));
    } else if (p == 1.0f || (std::isinf(p) && p > 0)) {
        if (param->rank() != 2) param = reshape(param, {param->shape()[0], param->volume() / param->shape()[0]});
        return matrixNorm(param, p);
    }
    throw Exception("Cannot compute convolution params norm for p=" + std::to_string(p));
}

Operation maxNorm(Operation param, Operation norm, Operation maxval) {
    return param * (1.0f / max(float32({}, {1.0f}), norm / maxval));
}

Projection projMatrix(Operation maxnorm, float p) {
    return [maxnorm, p](Operation param) { return maxNorm(param, matrixNorm(param, p), maxnorm); };
}
Projection projConvParams(Operation maxnorm, std::vector<size_t> inShape, std::vector<size_t> stride,
                          std::vector<size_t> padding, float p) {
    return [=](Operation param) { return maxNorm(param, convParamsNorm(param, inShape, stride, padding, p), maxnorm); };
}

Operation crossEntropy(Operation hypothesis, Operation groundTruth) {
    return sum(groundTruth * log(hypothesis + 1e-6f)) * (-1.0f / (float)hypothesis->shape()[0]);
}
Operation squaredError(Operation hypothesis, Operation groundTruth) {
    auto diff = hypothesis - groundTruth;
    return sum(diff * diff) * (1.0f / (float)hypothesis->shape()[0]);
}

// ---- nnet/models/vgg.d ------------------------------------------------------------------------------------------------------
LayerPtr vgg(Operation features, const std::vector<int>& sizes, std::vector<size_t> denseLayerSizes, VGGOptions opts) {
    auto layers = dataSource(features);
    for (int s : sizes) {   // makeExtractor, vgg.d:86-144
        if (s == -1) {
            layers = maxPool(layers, {2, 2});
        } else {
            Conv2DOptions co;
            co.padding = {1, 1};
            layers = conv2D(layers, (size_t)s, {3, 3}, co);
            if (opts.batchnorm) layers = batchNorm(layers);
            layers = relu(layers);
        }
    }
    for (auto s : denseLayerSizes) layers = relu(dense(layers, s));   // makeTop, vgg.d:146-174
    return layers;
}
LayerPtr vgg19(Operation features, std::vector<size_t> denseLayerSizes, VGGOptions opts) {
    return vgg(features, {64, 64, -1, 128, 128, -1, 256, 256, 256, 256, -1, 512, 512, 512, 512, -1, 512, 512, 512, 512, -1},
               denseLayerSizes, opts);
}

// ---- nnet/models/wrn.d ------------------------------------------------------------------------------------------------------
static LayerPtr meanPool(LayerPtr input) {
    // wrn.d:203-219: reshape -> sum([1]) (lowered to matmul with ones) -> reshape -> * (1/HW)
    auto impl = [](Operation inp) {
        size_t mapVol = inp->shape()[2] * inp->shape()[3];
        float scale = 1.0f / (float)mapVol;
        return reshape(sum(reshape(inp, {inp->shape()[0] * inp->shape()[1], mapVol}), {1}), {inp->shape()[0], inp->shape()[1]}) * scale;
    };
    auto y = impl(input->output());
    auto yTr = impl(input->trainOutput());
    return std::make_shared<Layer>(std::vector<LayerPtr>{input}, y, yTr, std::vector<Parameter>{});
}

static LayerPtr wrnBlock(LayerPtr inLayer, size_t u, size_t n, size_t s, const WRNOptions& opts) {
    // wrn.d:104-201
    auto convOpts = [&]() {
        Conv2DOptions o;
        o.padding = {1, 1};
        o.useBias = false;
        o.weightDecay = opts.weightDecay;
        return o;
    };
    LayerPtr res;
    for (size_t i = 0; i < n; ++i) {
        res = relu(batchNorm(inLayer));
        auto o1 = convOpts();
        o1.stride = {s, s};
        res = relu(batchNorm(conv2D(res, u, {3, 3}, o1)));
        res = conv2D(res, u, {3, 3}, convOpts());
        LayerPtr shortcut = inLayer;
        if (inLayer->output()->shape()[1] != res->output()->shape()[1]) {
            Conv2DOptions so;
            so.stride = {s, s};
            so.useBias = false;
            so.weightDecay = opts.weightDecay;
            shortcut = conv2D(inLayer, u, {1, 1}, so);
        }
        res = std::make_shared<Layer>(std::vector<LayerPtr>{res, shortcut}, res->output() + shortcut->output(),
                                      res->trainOutput() + shortcut->trainOutput(), std::vector<Parameter>{});
        inLayer = res;
        s = 1;
    }
    return res;
}

LayerPtr wideResNet(Operation features, size_t depth, size_t width, WRNOptions opts) {
    // wrn.d:56-102
    size_t n = (depth - 4) / 6;
    Conv2DOptions stem;
    stem.padding = {1, 1};
    stem.useBias = false;
    stem.weightDecay = opts.weightDecay;
    auto pred = conv2D(dataSource(features), 16, {3, 3}, stem);
    pred = wrnBlock(pred, 16 * width, n, opts.stride[0], opts);
    pred = wrnBlock(pred, 32 * width, n, opts.stride[1], opts);
    pred = wrnBlock(pred, 64 * width, n, opts.stride[2], opts);
    return meanPool(relu(batchNorm(pred)));
}

}  // namespace nnet
}  // namespace dopt
