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

// The max-gain projection shared by conv2D / dense / batchNorm (conv.d:125-150, dense.d:105-131, batchnorm.d:97-113):
// scale the new weights down so that the largest ratio ||after_n|| / ||before_n|| over the training batch stays below
// `maxgain`.  `before` / `after` are the layer's train-time input and output; they are flattened to [N, volume / N] here.
static Projection maxGainProjection(Operation before, Operation after, float maxgain, Projection inner) {
    before = reshape(before, {before->shape()[0], before->volume() / before->shape()[0]});
    after = reshape(after, {after->shape()[0], after->volume() / after->shape()[0]});
    return [before, after, maxgain, inner](Operation newWeights) {
        auto beforeNorms = sum(before * before, {1}) + 1e-8f;
        auto afterNorms = sum(after * after, {1}) + 1e-8f;
        auto mg = maxElement(sqrt(afterNorms / beforeNorms));
        auto projected = newWeights * (1.0f / max(float32Constant({}, {1.0f}), mg / maxgain));
        return inner ? inner(projected) : projected;
    };
}

// conv.d:173-189: the (deliberately simplified, Yoshida & Miyato 2017) spectral-norm penalty of the reshaped filter matrix --
// one power iteration from a random start; returns the SQUARED norm estimate, as the reference does
static Operation convSpectralNorm(Operation filters, size_t numIts = 1) {
    filters = reshape(filters, {filters->shape()[0], filters->volume() / filters->shape()[0]});
    auto x = uniformSample({filters->shape()[1], 1}) * 2.0f - 1.0f;
    for (size_t i = 0; i < numIts; ++i) x = matmul(transpose(filters, {1, 0}), matmul(filters, x));
    auto v = x / sqrt(sum(x * x));
    auto y = matmul(filters, v);
    return sum(y * y);
}

// dense.d:152-169: the same penalty for a dense layer's [outputs, inputs] weight matrix, power-iterating W W^T
static Operation denseSpectralNorm(Operation weights, size_t numIts = 1) {
    auto x = uniformSample({weights->shape()[0], 1}) * 2.0f - 1.0f;
    auto weightsT = transpose(weights, {1, 0});
    auto wwT = matmul(weights, weightsT);
    for (size_t i = 0; i < numIts; ++i) x = matmul(wwT, x);
    auto v = x / sqrt(sum(x * x));
    auto y = matmul(weightsT, v);
    return sum(y * y);
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
    filterLoss = safeAdd(filterLoss, opts.spectralDecay == 0.0f ? nullptr : (opts.spectralDecay * convSpectralNorm(filters)));
    auto y = convolution(x, filters, opts.padding, opts.stride);
    auto yTr = (xTr == x) ? y : convolution(xTr, filters, opts.padding, opts.stride);
    Projection filterProj = opts.filterProj;
    if (opts.maxgain != INFINITY) filterProj = maxGainProjection(xTr, yTr, opts.maxgain, opts.filterProj);   // conv.d:125-155
    std::vector<Parameter> params{Parameter{filters, filterLoss, filterProj}};
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
    weightLoss = safeAdd(weightLoss, opts.spectralDecay == 0.0f ? nullptr : (opts.spectralDecay * denseSpectralNorm(weights)));
    auto y = matmul(x, transpose(weights, {1, 0}));
    auto yTr = same ? y : matmul(xTr, transpose(weights, {1, 0}));
    Projection weightProj = opts.weightProj;
    if (opts.maxgain != INFINITY) weightProj = maxGainProjection(xTr, yTr, opts.maxgain, opts.weightProj);   // dense.d:105-136
    std::vector<Parameter> params{Parameter{weights, weightLoss, weightProj}};
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
    Projection gammaProj = opts.gammaProj;
    if (opts.maxgain != INFINITY) {
        // batchnorm.d:93-113: gain of the layer without its shift (zero beta, zero mean) on the train batch
        auto zeros = float32Constant({C}, std::vector<float>(C, 0.0f));
        auto after = batchNormInference(xTr, gamma, zeros, zeros, var);
        gammaProj = maxGainProjection(xTr, after, opts.maxgain, opts.gammaProj);
    } else if (opts.lipschitz != INFINITY) {
        // batchnorm.d:115-129
        float bound = opts.lipschitz;
        Projection inner = opts.gammaProj;
        gammaProj = [varUpdateSym, bound, inner](Operation newGamma) {
            auto norm = maxElement(abs(newGamma / sqrt(reshape(varUpdateSym, newGamma->shape()) + 1e-6f)));
            auto g = newGamma * (1.0f / max(float32Constant(1.0f), norm / bound));
            return inner ? inner(g) : g;
        };
    }
    std::vector<Parameter> params{
        Parameter{gamma, opts.gammaDecay == 0.0f ? nullptr : (opts.gammaDecay * sum(gamma * gamma)), gammaProj},
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
    bool ok = true;
    for (auto& p : mParams) {
        auto v = p->value()->get<float>();
        ok = ok && std::fwrite(v.data(), sizeof(float), v.size(), f) == v.size();
    }
    ok = (std::fclose(f) == 0) && ok;   // (a full disk must not leave a silently truncated parameter file behind)
    enforce(ok, "could not write " + filename);
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
void VGGOptions::verify() const {
    // vgg.d:24-42
    int regCtr = 0;
    if (!std::isnan(maxgainNorm)) {
        regCtr++;
        enforce(maxgainNorm == 2.0f, "Only a maxgainNorm of 2 is currently supported.");
    }
    if (!std::isnan(lipschitzNorm)) regCtr++;
    enforce(regCtr <= 1, "VGG models currently only support using one of maxgain and the lipschitz constraint");
}

LayerPtr vgg(Operation features, const std::vector<int>& sizes, std::vector<size_t> denseLayerSizes, VGGOptions opts) {
    opts.verify();
    const bool lip = !std::isnan(opts.lipschitzNorm);
    const float maxgain = std::isnan(opts.maxgainNorm) ? INFINITY : opts.maxNorm;
    auto layers = dataSource(features);
    int poolCtr = 0;
    for (int s : sizes) {   // makeExtractor, vgg.d:86-144
        if (s == -1) {
            layers = maxPool(layers, {2, 2});
            poolCtr++;
        } else {
            Conv2DOptions co;
            co.padding = {1, 1};
            co.maxgain = maxgain;
            co.spectralDecay = opts.spectralDecay;
            if (lip) {
                auto sh = layers->trainOutput()->shape();
                co.filterProj = projConvParams(float32Constant(opts.maxNorm), std::vector<size_t>(sh.begin() + 2, sh.end()),
                                               {1, 1}, {1, 1}, opts.lipschitzNorm);
            }
            if (opts.dropout && poolCtr != 0) layers = dropout(layers, 0.2f);
            layers = conv2D(layers, (size_t)s, {3, 3}, co);
            if (opts.batchnorm) {
                BatchNormOptions bo;
                bo.maxgain = maxgain;
                bo.lipschitz = lip ? opts.maxNorm : INFINITY;
                layers = batchNorm(layers, bo);
            }
            layers = relu(layers);
        }
    }
    for (auto s : denseLayerSizes) {   // makeTop, vgg.d:146-174
        DenseOptions d;
        d.maxgain = maxgain;
        d.spectralDecay = opts.spectralDecay;
        if (lip) d.weightProj = projMatrix(float32Constant(opts.maxNorm), opts.lipschitzNorm);
        if (opts.dropout) layers = dropout(layers, 0.5f);
        layers = relu(dense(layers, s, d));
    }
    return layers;
}
LayerPtr vgg16(Operation features, std::vector<size_t> denseLayerSizes, VGGOptions opts) {
    return vgg(features, {64, 64, -1, 128, 128, -1, 256, 256, 256, -1, 512, 512, 512, -1, 512, 512, 512, -1},
               denseLayerSizes, opts);
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

void WRNOptions::verify() const {
    // wrn.d:24-42
    int regCtr = 0;
    if (!std::isnan(maxgainNorm)) {
        regCtr++;
        enforce(maxgainNorm == 2.0f, "Only a maxgainNorm of 2 is currently supported.");
    }
    if (!std::isnan(lipschitzNorm)) regCtr++;
    enforce(regCtr <= 1, "VGG models currently only support using one of maxgain and the lipschitz constraint");
}

// the regulariser settings every layer of a WRN derives from its options (wrn.d:60-74,106-121)
struct WrnReg {
    float maxgain = INFINITY, lambda = INFINITY, lipschitzNorm = NAN;
    Operation lambdaSym;
    explicit WrnReg(const WRNOptions& o) {
        if (o.maxgainNorm == 2.0f) maxgain = o.maxNorm;
        if (!std::isnan(o.lipschitzNorm)) {
            lipschitzNorm = o.lipschitzNorm;
            lambda = o.maxNorm;
        }
        lambdaSym = float32Constant(lambda);
    }
    BatchNormOptions bnOpts() const {
        BatchNormOptions b;
        b.maxgain = maxgain;
        b.lipschitz = lambda;
        return b;
    }
    // operator-norm projection for a convolution reading `in` (null when the constraint is off)
    Projection proj(const LayerPtr& in, size_t stride, size_t pad) const {
        if (lambda == INFINITY) return nullptr;
        auto sh = in->trainOutput()->shape();
        return projConvParams(lambdaSym, std::vector<size_t>(sh.begin() + 2, sh.end()), {stride, stride}, {pad, pad},
                              lipschitzNorm);
    }
};

static LayerPtr wrnBlock(LayerPtr inLayer, size_t u, size_t n, size_t s, const WRNOptions& opts) {
    // wrn.d:104-201
    WrnReg reg(opts);
    auto convOpts = [&]() {
        Conv2DOptions o;
        o.padding = {1, 1};
        o.useBias = false;
        o.weightDecay = opts.weightDecay;
        o.spectralDecay = opts.spectralDecay;
        o.maxgain = reg.maxgain;
        return o;
    };
    LayerPtr res;
    for (size_t i = 0; i < n; ++i) {
        res = relu(batchNorm(inLayer, reg.bnOpts()));
        auto o1 = convOpts();
        o1.stride = {s, s};
        // NB the reference passes padding [1, 1] to every projConvParams call, the 1x1 shortcut included (wrn.d:146,164,177)
        o1.filterProj = reg.proj(res, s, 1);
        res = relu(batchNorm(conv2D(res, u, {3, 3}, o1), reg.bnOpts()));
        if (opts.dropout) res = dropout(res, 0.3f);   // maybeDropout, wrn.d:159
        auto o2 = convOpts();
        o2.filterProj = reg.proj(res, 1, 1);
        res = conv2D(res, u, {3, 3}, o2);
        LayerPtr shortcut = inLayer;
        if (inLayer->output()->shape()[1] != res->output()->shape()[1]) {
            Conv2DOptions so;
            so.stride = {s, s};
            so.useBias = false;
            so.weightDecay = opts.weightDecay;
            so.spectralDecay = opts.spectralDecay;
            so.maxgain = reg.maxgain;
            so.filterProj = reg.proj(inLayer, s, 1);
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
    opts.verify();
    WrnReg reg(opts);
    auto src = dataSource(features);
    Conv2DOptions stem;
    stem.padding = {1, 1};
    stem.useBias = false;
    stem.weightDecay = opts.weightDecay;
    stem.spectralDecay = opts.spectralDecay;
    stem.maxgain = reg.maxgain;
    stem.filterProj = reg.proj(src, 1, 1);
    auto pred = conv2D(src, 16, {3, 3}, stem);
    pred = wrnBlock(pred, 16 * width, n, opts.stride[0], opts);
    pred = wrnBlock(pred, 32 * width, n, opts.stride[1], opts);
    pred = wrnBlock(pred, 64 * width, n, opts.stride[2], opts);
    return meanPool(relu(batchNorm(pred, reg.bnOpts())));
}

}  // namespace nnet
}  // namespace dopt
