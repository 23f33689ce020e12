// Drop-in for include/gridStructure.h:33-58 (GridStructure, GridWindow, getLineCoords): host-side containers that
// Frame::ComputeStereoMatches_Lines fills (src/Frame.cc:896-913) and hands to matchGrid.  Inside the reference tree the
// reference's own gridStructure.cpp / LineIterator.cpp (OpenCV-free, 160 lines) stay as they are; this header is the
// stand-in for builds outside it.  Storage is one flat vector of cells; the API is the reference's.
#pragma once
#include <list>
#include <stdexcept>
#include <unordered_set>
#include <utility>
#include <vector>
namespace ORB_SLAM2 {
struct GridWindow { std::pair<int, int> width, height; };
class GridStructure {
public:
    int rows, cols;
    GridStructure(int rows_, int cols_) : rows(rows_), cols(cols_) {
        if (rows <= 0 || cols <= 0) throw std::runtime_error("[GridStructure] invalid dimension");
        cells_.resize((size_t)rows * cols);
    }
    std::list<int>& at(int x, int y) { return inside(x, y) ? cells_[(size_t)x * rows + y] : out_of_bounds_; }
    const std::list<int>& cell(int x, int y) const { return cells_[(size_t)x * rows + y]; }
    void get(int x, int y, const GridWindow& w, std::unordered_set<int>& indices) const {
        const int x0 = std::max(0, x - w.width.first), x1 = std::min(cols, x + w.width.second + 1);
        const int y0 = std::max(0, y - w.height.first), y1 = std::min(rows, y + w.height.second + 1);
        for (int cx = x0; cx < x1; ++cx)
            for (int cy = y0; cy < y1; ++cy) indices.insert(cell(cx, cy).begin(), cell(cx, cy).end());
    }
    void clear() { for (auto& c : cells_) c.clear(); }
private:
    bool inside(int x, int y) const { return x >= 0 && x < cols && y >= 0 && y < rows; }
    std::vector<std::list<int>> cells_;
    std::list<int> out_of_bounds_;
};
// Bresenham walk of src/LineIterator.cpp:34-77 (double coordinates, steep / swap handling) into (x, y) cells
inline void getLineCoords(double x1, double y1, double x2, double y2, std::list<std::pair<int, int>>& line_coords) {
    line_coords.clear();
    const bool steep = std::abs(y2 - y1) > std::abs(x2 - x1);
    if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
    if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
    const double dx = x2 - x1, dy = std::abs(y2 - y1);
    double error = dx / 2.0;
    const int ystep = (y1 < y2) ? 1 : -1;
    int y = (int)y1;
    const int maxX = (int)x2;
    for (int x = (int)x1; x <= maxX; ++x) {
        line_coords.push_back(steep ? std::make_pair(y, x) : std::make_pair(x, y));
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
    }
}
}  // namespace ORB_SLAM2
