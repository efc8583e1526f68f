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
  // OpenCV 3: compute() is detectAndCompute() on the provided key points
  virtual void compute(InputArray image, std::vector<KeyPoint>& k, OutputArray descriptors) {
    detectAndCompute(image, noArray(), k, descriptors, true);
  }
  virtual int descriptorSize() const { return 0; }
  virtual int descriptorType() const { return 0; }
};

// cv::DMatch (features2d.hpp): ordered by distance only.
struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(std::numeric_limits<float>::max()) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  bool operator<(const DMatch& m) const { return distance < m.distance; }
};

// The slice of cv::DescriptorMatcher that brisk::BruteForceMatcher builds on (OpenCV 3 features2d
// matchers.cpp): the train collection, the public knnMatch / radiusMatch entry points that forward to the
// protected *Impl virtuals, and the two mask helpers.
class DescriptorMatcher {
 public:
  virtual ~DescriptorMatcher() {}
  virtual void add(const std::vector<Mat>& descriptors) {
    trainDescCollection.insert(trainDescCollection.end(), descriptors.begin(), descriptors.end());
  }
  const std::vector<Mat>& getTrainDescriptors() const { return trainDescCollection; }
  virtual void clear() { trainDescCollection.clear(); }
  virtual bool empty() const { return trainDescCollection.empty(); }
  virtual bool isMaskSupported() const = 0;
  virtual void train() {}
  virtual Ptr<DescriptorMatcher> clone(bool emptyTrainData = false) const = 0;
  void knnMatch(InputArray query, std::vector<std::vector<DMatch> >& matches, int k, InputArrayOfArrays masks = noArray(),
                bool compactResult = false) {
    matches.clear();
    if (empty() || query.empty()) return;
    train();
    knnMatchImpl(query, matches, k, masks, compactResult);
  }
  void radiusMatch(InputArray query, std::vector<std::vector<DMatch> >& matches, float maxDistance,
                   InputArrayOfArrays masks = noArray(), bool compactResult = false) {
    matches.clear();
    if (empty() || query.empty()) return;
    train();
    radiusMatchImpl(query, matches, maxDistance, masks, compactResult);
  }

 protected:
  virtual void knnMatchImpl(InputArray query, std::vector<std::vector<DMatch> >& matches, int k,
                            InputArrayOfArrays masks = noArray(), bool compactResult = false) = 0;
  virtual void radiusMatchImpl(InputArray query, std::vector<std::vector<DMatch> >& matches, float maxDistance,
                               InputArrayOfArrays masks = noArray(), bool compactResult = false) = 0;
  static bool isPossibleMatch(const Mat& mask, int queryIdx, int trainIdx) {
    return mask.empty() || mask.at<unsigned char>(queryIdx, trainIdx);
  }
  // every mask non-empty with an all-zero row queryIdx (an empty mask never counts)
  static bool isMaskedOut(const std::vector<Mat>& masks, int queryIdx) {
    size_t outCount = 0;
    for (size_t i = 0; i < masks.size(); i++)
      if (!masks[i].empty() && countNonZero(masks[i].row(queryIdx)) == 0) outCount++;
    return !masks.empty() && outCount == masks.size();
  }
  static Mat clone_op(Mat m) { return m.clone(); }
  std::vector<Mat> trainDescCollection;
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;
}  // namespace cv
