// ORACLE SCAFFOLDING (test infrastructure): minimal cv::Feature2D base.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
class Feature2D {
 public:
  virtual ~Feature2D() {}
  virtual void detectAndCompute(InputArray, InputArray, std::vector<KeyPoint>&, OutputArray, bool = false) {}
  virtual void detect(InputArray image, std::vector<KeyPoint>& k, InputArray mask = noArray()) {
    detectAndCompute(image, mask, k, noOutArray(), false);
  }
  virtual int descriptorSize() const { return 0; }
  virtual int descriptorType() const { return 0; }
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;
}  // namespace cv
