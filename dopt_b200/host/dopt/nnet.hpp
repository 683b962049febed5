// dopt/nnet.hpp -- C++ mirror of the parts of dopt.nnet that generate the hot path's graphs: Layer, the layer
// constructors, DAGNetwork, the losses and the VGG / Wide-ResNet model builders.  Host-only graph construction; see the
// .cpp for per-function citations.  The research regularisers (nnet/lipschitz.d projections, the per-layer maxgain /
// spectralDecay / lipschitz options, dropout) are default-off in every BASELINE config but mirrored.
#pragma once
#include <cmath>
#include <random>

#include "core.hpp"

namespace dopt {
namespace nnet {

using Projection = std::function<Operation(Operation)>;   // online/package.d:28

// nnet/parameters.d: a symbol, an optional loss term and an optional projection
struct Parameter {
    Operation symbol;
    Operation loss;
    Projection projection;
};
using ParamInitializer = std::function<void(Operation)>;
ParamInitializer constantInit(float v);    // parameters.d:86-102
ParamInitializer heGaussianInit();         // parameters.d:262-272: N(0, sqrt(2 / fanIn)), fanIn = prod(shape[1..])
void seedInitializers(uint64_t seed);      // the reference draws from an unseeded std.random; tests need determinism

class Layer;
using LayerPtr = std::shared_ptr<Layer>;
class Layer {   // nnet/layers/package.d:30-76
public:
    Layer(std::vector<LayerPtr> deps, Operation outExpr, Operation trainOutExpr, std::vector<Parameter> params)
        : mDeps(std::move(deps)), mParams(std::move(params)), mOutput(outExpr), mTrainOutput(trainOutExpr) {}
    const std::vector<LayerPtr>& deps() const { return mDeps; }
    const std::vector<Parameter>& params() const { return mParams; }
    Operation output() const { return mOutput; }
    Operation trainOutput() const { return mTrainOutput; }
private:
    std::vector<LayerPtr> mDeps;
    std::vector<Parameter> mParams;
    Operation mOutput, mTrainOutput;
};
std::vector<LayerPtr> topologicalSort(const std::vector<LayerPtr>& layers);

struct Conv2DOptions {   // nnet/layers/conv.d:16-43
    std::vector<size_t> padding{0, 0}, stride{1, 1};
    ParamInitializer filterInit = heGaussianInit(), biasInit = constantInit(0.0f);
    Projection filterProj, biasProj;
    float weightDecay = 0.0f;
    float maxgain = INFINITY;       // conv.d:27,128-150: project so that max_n ||y_n|| / ||x_n|| <= maxgain on the train batch
    float spectralDecay = 0.0f;     // conv.d:28,112-115: + spectralDecay * (squared spectral-norm estimate of the filters)
    bool useBias = true;
};
struct DenseOptions {    // nnet/layers/dense.d:15-37
    ParamInitializer weightInit = heGaussianInit(), biasInit = constantInit(0.0f);
    Projection weightProj, biasProj;
    float weightDecay = 0.0f;
    float maxgain = INFINITY;       // dense.d:108-131
    float spectralDecay = 0.0f;     // dense.d:99-102
    bool useBias = true;
};
struct BatchNormOptions {   // nnet/layers/batchnorm.d:14-38
    ParamInitializer gammaInit = constantInit(1.0f), betaInit = constantInit(0.0f);
    Projection gammaProj, betaProj;
    float gammaDecay = 0.0f;
    float momentum = 0.9f;
    float maxgain = INFINITY;       // batchnorm.d:97-113
    float lipschitz = INFINITY;     // batchnorm.d:115-129: bound max_c |gamma_c| / sqrt(var_c + 1e-6)
};

LayerPtr dataSource(Operation var);
LayerPtr dataSource(Operation var, Operation trainVar);
LayerPtr conv2D(LayerPtr input, size_t outputChannels, std::vector<size_t> filterDims, Conv2DOptions opts = Conv2DOptions());
LayerPtr dense(LayerPtr input, size_t numOutputs, DenseOptions opts = DenseOptions());
LayerPtr batchNorm(LayerPtr input, BatchNormOptions opts = BatchNormOptions());
LayerPtr relu(LayerPtr input);
LayerPtr maxPool(LayerPtr input, std::vector<size_t> dims);
LayerPtr dropout(LayerPtr input, float dropProb);   // nnet/layers/dropout.d:14-29
LayerPtr softmax(LayerPtr input);

class DAGNetwork {   // nnet/networks.d:24-128
public:
    DAGNetwork(std::vector<Operation> inputs, std::vector<LayerPtr> outputs);
    const std::vector<Operation>& inputs() const { return mInputs; }
    const std::vector<Operation>& outputs() const { return mOutputs; }
    const std::vector<Operation>& trainOutputs() const { return mTrainOutputs; }
    Operation paramLoss() const { return mParameterLoss; }
    const std::map<Operation, Projection>& paramProj() const { return mParameterProj; }
    const std::vector<Operation>& params() const { return mParams; }
    // networks.d:130-164: raw fp32 of every parameter in order, no header
    void save(const std::string& filename) const;
    void load(const std::string& filename);
private:
    std::vector<Operation> mInputs, mOutputs, mTrainOutputs, mParams;
    Operation mParameterLoss;
    std::map<Operation, Projection> mParameterProj;
};

// nnet/lipschitz.d (Gouk et al. 2018): projections that bound the operator norm of a weight matrix / convolution, for the
// `projs` argument of the dopt.online updaters.  p is 1, 2 (n power iterations from a random start) or infinity.
Operation matrixNorm(Operation param, float p, size_t n = 2);                                   // lipschitz.d:43-97
Operation convParamsNorm(Operation param, std::vector<size_t> inShape, std::vector<size_t> stride,
                         std::vector<size_t> padding, float p = 2.0f, size_t n = 2);            // lipschitz.d:111-147
Operation maxNorm(Operation param, Operation norm, Operation maxval);                           // lipschitz.d:162-165
Projection projMatrix(Operation maxnorm, float p = 2.0f);                                       // lipschitz.d:28-38
Projection projConvParams(Operation maxnorm, std::vector<size_t> inShape, std::vector<size_t> stride,
                          std::vector<size_t> padding, float p = 2.0f);                         // lipschitz.d:99-109

Operation crossEntropy(Operation hypothesis, Operation groundTruth);   // nnet/losses.d:23-26
Operation squaredError(Operation hypothesis, Operation groundTruth);   // nnet/losses.d:35-40

struct VGGOptions {   // nnet/models/vgg.d:12-49
    bool dropout = false;           // 0.2 before every convolution after the first pool, 0.5 before every dense layer
    bool batchnorm = false;
    float maxgainNorm = NAN;        // only 2 is supported
    float lipschitzNorm = NAN;
    float maxNorm = INFINITY;
    float spectralDecay = 0.0f;
    void verify() const;
};
LayerPtr vgg16(Operation features, std::vector<size_t> denseLayerSizes = {4096, 4096}, VGGOptions opts = VGGOptions());
LayerPtr vgg19(Operation features, std::vector<size_t> denseLayerSizes = {4096, 4096}, VGGOptions opts = VGGOptions());
LayerPtr vgg(Operation features, const std::vector<int>& extractorSizes, std::vector<size_t> denseLayerSizes,
             VGGOptions opts = VGGOptions());

struct WRNOptions {   // nnet/models/wrn.d:11-54
    bool dropout = false;           // 0.3 after the first conv-bn-relu of every block (wrn.d:159)
    float maxgainNorm = NAN;        // only 2 is supported (wrn.d:32)
    float lipschitzNorm = NAN;      // 1, 2 or infinity: operator-norm constraint on every convolution (wrn.d:118-122)
    float maxNorm = INFINITY;       // the bound used by whichever of the two is enabled
    float spectralDecay = 0.0f;
    float weightDecay = 0.0001f;
    size_t stride[3] = {1, 2, 2};
    void verify() const;            // wrn.d:24-42
};
LayerPtr wideResNet(Operation features, size_t depth, size_t width, WRNOptions opts = WRNOptions());

}  // namespace nnet
}  // namespace dopt
