// introsort.h -- exact restatement of libstdc++'s std::sort (bits/stl_algo.h:
// __introsort_loop / __unguarded_partition_pivot / __move_median_to_first /
// __final_insertion_sort, heap fallback from bits/stl_heap.h).
//
// Why: FeatureTracker::setMask orders the tracked features with an *unstable*
// std::sort on track_cnt only (reference vins_estimator/src/feature_tracker/
// feature_tracker.cpp:186-188).  The tie order decides which feature wins the
// greedy min-distance mask and therefore the feature IDs, so the product has to
// reproduce libstdc++'s exact move sequence.  The element is (key, payload);
// only the key is compared (descending: comp(a,b) := a.key > b.key).
//
// Usable from host and device code (single thread).
#pragma once

#ifdef __CUDACC__
#define VRF_HD __host__ __device__ __forceinline__
#else
#define VRF_HD inline
#endif

namespace vrf {

struct SortItem {
    int key;   // track_cnt
    int val;   // original index
};

VRF_HD bool si_comp(const SortItem &a, const SortItem &b) { return a.key > b.key; }
VRF_HD void si_swap(SortItem &a, SortItem &b) { SortItem t = a; a = b; b = t; }

// bits/stl_heap.h
VRF_HD void si_push_heap(SortItem *first, int hole, int top, SortItem value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && si_comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

VRF_HD void si_adjust_heap(SortItem *first, int hole, int len, SortItem value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (si_comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    si_push_heap(first, hole, top, value);
}

VRF_HD void si_heap_sort(SortItem *first, int len)   // std::__partial_sort(first,last,last)
{
    if (len >= 2) {   // __make_heap
        int parent = (len - 2) / 2;
        while (true) {
            SortItem v = first[parent];
            si_adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;   // __sort_heap
    while (last > 1) {
        --last;
        SortItem v = first[last];     // __pop_heap(first, last, last)
        first[last] = first[0];
        si_adjust_heap(first, 0, last, v);
    }
}

VRF_HD void si_move_median_to_first(SortItem *a_, int result, int a, int b, int c)
{
    if (si_comp(a_[a], a_[b])) {
        if (si_comp(a_[b], a_[c])) si_swap(a_[result], a_[b]);
        else if (si_comp(a_[a], a_[c])) si_swap(a_[result], a_[c]);
        else si_swap(a_[result], a_[a]);
    } else if (si_comp(a_[a], a_[c])) si_swap(a_[result], a_[a]);
    else if (si_comp(a_[b], a_[c])) si_swap(a_[result], a_[c]);
    else si_swap(a_[result], a_[b]);
}

VRF_HD int si_unguarded_partition(SortItem *a, int first, int last, int pivot)
{
    while (true) {
        while (si_comp(a[first], a[pivot])) ++first;
        --last;
        while (si_comp(a[pivot], a[last])) --last;
        if (!(first < last)) return first;
        si_swap(a[first], a[last]);
        ++first;
    }
}

VRF_HD void si_unguarded_linear_insert(SortItem *a, int last)
{
    SortItem val = a[last];
    int next = last - 1;
    while (si_comp(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

VRF_HD void si_insertion_sort(SortItem *a, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (si_comp(a[i], a[first])) {
            SortItem val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];   // move_backward
            a[first] = val;
        } else
            si_unguarded_linear_insert(a, i);
    }
}

// std::sort(a, a+n, comp) with comp = descending key.  Explicit stack instead
// of recursion (device friendly): the right part is "recursed" first exactly as
// libstdc++ does (__introsort_loop(cut, last, depth); last = cut).
VRF_HD void std_sort_desc(SortItem *a, int n)
{
    if (n <= 1) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;           // std::__lg(n)
    // frames: (first, last, depth_limit)
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        // __introsort_loop body.  The recursive call on [cut,last) runs to
        // completion BEFORE the loop continues on [first,cut).  Both touch
        // disjoint ranges, so processing order does not change the result;
        // we push the left continuation and then handle the right part first
        // to keep the same order anyway.
        while (last - first > 16) {
            if (depth == 0) {
                si_heap_sort(a + first, last - first);
                last = first;   // done with this range
                break;
            }
            --depth;
            int mid = first + (last - first) / 2;
            si_move_median_to_first(a, first, first + 1, mid, last - 1);
            int cut = si_unguarded_partition(a, first + 1, last, first);
            // recurse on [cut,last) now; continue with [first,cut) afterwards
            stk_first[sp] = first; stk_last[sp] = cut; stk_depth[sp] = depth; sp++;
            first = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        si_insertion_sort(a, 0, 16);
        for (int i = 16; i != n; ++i) si_unguarded_linear_insert(a, i);
    } else
        si_insertion_sort(a, 0, n);
}

}  // namespace vrf
