#!/usr/bin/env python3
"""ORACLE -- TEST INFRASTRUCTURE ONLY.
Cuts function definitions, verbatim, out of the reference's large translation units (which as a whole need OpenCV / Eigen /
Pangolin and cannot be compiled here) into oracle/_ref/gen/*.inc, so that the harness can compile the reference's OWN text of
those functions against the stand-in types of cvstub.hpp / frame_stub.hpp.  Nothing is rewritten: a slice is the source
lines from the line that starts with the given signature to the brace that closes the definition.  The output directory is
git-ignored (reference sources are never committed)."""
import pathlib, re, sys

REF = pathlib.Path(sys.argv[1]); OUT = pathlib.Path(sys.argv[2]); OUT.mkdir(parents=True, exist_ok=True)


def cut(path, start, end_semicolon=False):
    lines = (REF / path).read_text(errors="replace").splitlines()
    for i, ln in enumerate(lines):
        if ln.startswith(start):
            break
    else:
        raise SystemExit(f"slice.py: '{start}' not found in {path}")
    depth, seen, out = 0, False, []
    for ln in lines[i:]:
        out.append(ln)
        code = re.sub(r'//.*', '', ln)
        depth += code.count("{") - code.count("}")
        seen = seen or "{" in code
        if seen and depth == 0 and (not end_semicolon or code.rstrip().endswith(";")):
            break
    return f"// ---- {path}:{i + 1}-{i + len(out)} (verbatim) ----\n" + "\n".join(out) + "\n"


LBD = "Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp"
SLICES = {
    "lbd.inc": [(LBD, "static const int combinations[32][2] =", True), (LBD, "BinaryDescriptor::Params::Params()"),
                (LBD, "BinaryDescriptor::BinaryDescriptor( const BinaryDescriptor::Params &parameters ) :"),
                (LBD, "static inline int get2Pow( int i )"), (LBD, "void BinaryDescriptor::computeGaussianPyramid("),
                (LBD, "void BinaryDescriptor::computeSobel("), (LBD, "unsigned char BinaryDescriptor::binaryConversion("),
                (LBD, "void BinaryDescriptor::computeImpl("), (LBD, "int BinaryDescriptor::computeLBD(")],
    "orbmatcher.inc": [("src/ORBmatcher.cc", "int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)"),
                       ("src/ORBmatcher.cc", "float ORBmatcher::RadiusByViewingCos("),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono, map<int, int>& match12)"),
                       ("src/ORBmatcher.cc", "void ORBmatcher::ComputeThreeMaxima("), ("src/ORBmatcher.cc", "int ORBmatcher::DescriptorDistance("),
                       # SURVEY 8f rank 2
                       ("src/ORBmatcher.cc", "bool ORBmatcher::CheckDistEpipolarLine("),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw,"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2,"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F,"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchForTriangulation("),
                       ("src/ORBmatcher.cc", "int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw,"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchBySim3("),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF,"),
                       ("src/ORBmatcher.cc", "int ORBmatcher::SearchForInitialization(")],
    "mappoint.inc": [("src/MapPoint.cc", "void MapPoint::ComputeDistinctiveDescriptors()"),
                     ("src/MapPoint.cc", "float MapPoint::GetMinDistanceInvariance()"), ("src/MapPoint.cc", "float MapPoint::GetMaxDistanceInvariance()"),
                     ("src/MapPoint.cc", "int MapPoint::PredictScale(const float &currentDist, KeyFrame* pKF)"),
                     ("src/MapPoint.cc", "int MapPoint::PredictScale(const float &currentDist, Frame* pF)")],
    "keyframe.inc": [("src/KeyFrame.cc", "vector<size_t> KeyFrame::GetFeaturesInArea("), ("src/KeyFrame.cc", "bool KeyFrame::IsInImage(")],
    "frame.inc": [("src/Frame.cc", "void Frame::AssignFeaturesToGrid()"), ("src/Frame.cc", "vector<size_t> Frame::GetFeaturesInArea("),
                  ("src/Frame.cc", "bool Frame::PosInGrid("), ("src/Frame.cc", "void Frame::ComputeStereoMatches()"),
                  ("src/Frame.cc", "void Frame::ComputeStereoMatches_Lines("), ("src/Frame.cc", "double Frame::lineSegmentOverlapStereo("),
                  ("src/Frame.cc", "void Frame::filterLineSegmentDisparity(")],
}
for name, parts in SLICES.items():
    (OUT / name).write_text("".join(cut(*p) for p in parts))
print("sliced", ", ".join(SLICES))
