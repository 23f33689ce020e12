// The reference's data classes as the shim sources see them.  Inside the reference tree (-DOLF_IN_REFERENCE_TREE) they ARE the
// reference's own headers (include/Frame.h, MapPoint.h, MapLine.h, KeyFrame.h, ORBmatcher.h, gridStructure.h -- no file of this
// directory carries one of those names, so the quoted includes below resolve to the reference's include/); elsewhere the
// stand-ins of shim/standin/ with the same members.
#pragma once
#ifdef OLF_IN_REFERENCE_TREE
#include "Frame.h"
#include "MapPoint.h"
#include "MapLine.h"
#include "KeyFrame.h"
#include "ORBmatcher.h"
#include "gridStructure.h"
namespace ORB_SLAM2 { inline void olf_set_le(Frame& F, size_t i, double a, double b, double c) { F.mvle_l[i] = Vector3d(a, b, c); } }
#else
#include "standin/Frame.h"
#include "standin/ORBmatcher.h"
#include "standin/gridStructure.h"
#endif
