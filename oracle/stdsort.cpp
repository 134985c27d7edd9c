// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// The reference orders tracked features with libstdc++'s *unstable*
// std::sort (vins_estimator/src/feature_tracker/feature_tracker.cpp:186-188,
// comparator `a.first > b.first` on pair<int, pair<Point2f,int>>), so the tie
// order of equal track counts is libstdc++-introsort specific and feeds the
// greedy mask and therefore the feature IDs.  This helper runs the *real*
// std::sort of this image's libstdc++ (g++ 13) with the same comparator on
// the same element shape, so the python oracle reproduces the reference's tie
// order exactly.  Element payload does not influence std::sort's moves.
#include <algorithm>
#include <utility>
#include <vector>

struct P2f { float x, y; };

extern "C" void oracle_stdsort_desc_by_cnt(const int *cnt, int n, int *perm_out)
{
    std::vector<std::pair<int, std::pair<P2f, int>>> v;
    v.reserve(n);
    for (int i = 0; i < n; i++)
        v.emplace_back(cnt[i], std::make_pair(P2f{0.f, 0.f}, i));
    std::sort(v.begin(), v.end(),
              [](const std::pair<int, std::pair<P2f, int>> &a,
                 const std::pair<int, std::pair<P2f, int>> &b) { return a.first > b.first; });
    for (int i = 0; i < n; i++)
        perm_out[i] = v[i].second.second;
}
