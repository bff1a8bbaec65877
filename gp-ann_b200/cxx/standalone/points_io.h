// points_io.h (standalone builds only) -- gp-ann's src/points_io.h:5-11 with the same signatures and file formats.
// Inside a gp-ann checkout the reference's own points_io.{h,cpp} is used.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "defs.h"

PointSet ReadPoints(const std::string& path, int64_t size = -1);
void WritePoints(PointSet& points, const std::string& path);
std::vector<NNVec> ReadGroundTruth(const std::string& path);
void WriteGroundTruth(const std::string& path, const std::vector<NNVec>& ground_truth);
