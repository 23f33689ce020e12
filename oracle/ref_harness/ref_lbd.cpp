// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// The LBD functions of Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp (verbatim slices, see slice.py) in a
// translation unit of their own that starts exactly like the reference's: `#include "precomp_custom.hpp"` (:42).  That header
// pulls in bitarray_custom.hpp -> <math.h>, whose C++ wrapper (libstdc++ >= 6) puts the float overloads into the global
// namespace: unqualified cos / sin of a float argument are cosf / sinf in the reference's build, and this one.
#include "precomp_custom.hpp"
namespace cv { namespace line_descriptor {
#define NUM_OF_BANDS 9
#include "lbd.inc"
// members of the reference class that are declared virtual (or constructed) but not on the hot path: never called here
BinaryDescriptor::~BinaryDescriptor() {}
void BinaryDescriptor::read(const cv::FileNode&) {}
void BinaryDescriptor::write(cv::FileStorage&) const {}
void BinaryDescriptor::operator()(InputArray, InputArray, std::vector<KeyLine>&, OutputArray, bool, bool) const { abort(); }
void BinaryDescriptor::detectImpl(const Mat&, std::vector<KeyLine>&, const Mat&) const { abort(); }
BinaryDescriptor::EDLineDetector::EDLineDetector() {}
BinaryDescriptor::EDLineDetector::~EDLineDetector() {}
Ptr<BinaryDescriptor> BinaryDescriptor::createBinaryDescriptor() { return Ptr<BinaryDescriptor>(new BinaryDescriptor()); }     // (:205-208)
void BinaryDescriptor::compute(const Mat& image, std::vector<KeyLine>& keylines, Mat& descriptors, bool returnFloatDescr) const {   // (:524-528)
    computeImpl(image, keylines, descriptors, returnFloatDescr, false);
}
} }
