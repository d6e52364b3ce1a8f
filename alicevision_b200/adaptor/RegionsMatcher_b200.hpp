// Reference-side adaptor, Surface 1b: a matching::IRegionsMatcher (matching/RegionsMatcher.hpp:49-78) whose Match runs the
// whole RegionsMatcher<ArrayMatcherT>::Match body (:126-176: top-2 search, ratio test, both de-duplications) on the B200
// engine, i.e. on the tensor-core kernel whenever the descriptors qualify.  This is what the factory
// createRegionsMatcher (matching/RegionsMatcher.cpp:54-176) returns for the new enum values, so every caller of
// RegionsDatabaseMatcher / DistanceRatioMatch (:19-52; localization/VoctreeLocalizer.cpp:399,763) gets the GPU path.
// Header-only; compiled inside an AliceVision build and linked against libb200match.so.
#pragma once

#include <aliceVision/matching/RegionsMatcher.hpp>
#include <aliceVision/feature/Regions.hpp>

#include <b200match.h>

#include "ArrayMatcher_b200.hpp"   // b200detail::sharedContext (same directory: matching/)

#include <atomic>
#include <memory>
#include <typeinfo>
#include <vector>

namespace aliceVision {
namespace matching {

namespace b200detail {
/// View ids of the shared context handed out to IRegionsMatcher instances (database and per-call query views).
inline uint32_t nextViewId()
{
    static std::atomic<uint32_t> next{0x40000000u};
    return next.fetch_add(1u);
}

/// Element type of a Regions as the engine names it, or -1 when createRegionsMatcher has no brute-force case for it.
inline int regionsDtype(const feature::Regions& r)
{
    if (r.IsBinary())
        return r.Type_id() == typeid(unsigned char).name() ? B200M_BIN : -1;
    if (r.Type_id() == typeid(float).name())
        return B200M_F32;
    if (r.Type_id() == typeid(unsigned char).name())
        return B200M_U8;
    return -1;
}

inline bool uploadRegions(b200m_ctx* ctx, uint32_t id, const feature::Regions& r, int dtype)
{
    const int n = static_cast<int>(r.RegionCount());
    std::vector<float> xy(2 * static_cast<size_t>(n));
    const auto& feats = r.Features();
    for (int k = 0; k < n; ++k)
    {
        xy[2 * k] = feats[k].x();
        xy[2 * k + 1] = feats[k].y();
    }
    return b200m_upload_view(ctx, id, n ? r.DescriptorRawData() : nullptr, n, static_cast<int>(r.DescriptorLength()), dtype, xy.data()) == B200M_OK;
}
}  // namespace b200detail

class RegionsMatcher_b200 : public IRegionsMatcher
{
  public:
    /// Same arguments as RegionsMatcher<ArrayMatcherT>(rng, regions, b_squared_metric) (:105-114); the database descriptors
    /// and positions are copied to the device once, here.  b_squared_metric is implied by the element type, as in the
    /// factory (true for the L2 cases, false for Hamming: RegionsMatcher.cpp:78,107,136,163).
    RegionsMatcher_b200(std::mt19937& /*randomNumberGenerator*/, const feature::Regions& regions)
      : IRegionsMatcher(regions),
        _ctx(b200detail::sharedContext()),
        _dtype(b200detail::regionsDtype(regions))
    {
        if (_ctx == nullptr || _dtype < 0 || regions.RegionCount() == 0)
            return;   // like an ArrayMatcher whose Build failed: Match returns false
        _dbId = b200detail::nextViewId();
        _built = b200detail::uploadRegions(_ctx, _dbId, regions, _dtype);
    }
    ~RegionsMatcher_b200() override
    {
        if (_built)
            b200m_remove_view(_ctx, _dbId);
    }
    RegionsMatcher_b200(const RegionsMatcher_b200&) = delete;
    RegionsMatcher_b200& operator=(const RegionsMatcher_b200&) = delete;

    /// RegionsMatcher::Match (:126-176): false when the query is empty, of another type, or the search fails;
    /// otherwise the de-duplicated IndMatch(i = database feature, j = query feature) list and `!empty()`.
    bool Match(const float f_dist_ratio, const feature::Regions& query_regions, matching::IndMatches& vec_putative_matches) override
    {
        if (query_regions.RegionCount() == 0)
            return false;
        if (!_built || b200detail::regionsDtype(query_regions) != _dtype || query_regions.DescriptorLength() != this->regions_.DescriptorLength())
            return false;
        const uint32_t qId = b200detail::nextViewId();
        if (!b200detail::uploadRegions(_ctx, qId, query_regions, _dtype))
            return false;
        const uint32_t pair[2] = {_dbId, qId};
        b200m_result* res = nullptr;
        const int rc = b200m_match_pairs(_ctx, pair, 1, f_dist_ratio, 0, B200M_STAGE_FULL, &res);
        if (rc == B200M_OK)
        {
            const int64_t* off = nullptr;
            const b200m_match* m = nullptr;
            b200m_result_get(res, nullptr, &off, &m);
            vec_putative_matches.reserve(vec_putative_matches.size() + static_cast<size_t>(off[1]));
            for (int64_t e = off[0]; e < off[1]; ++e)
                vec_putative_matches.emplace_back(m[e].i, m[e].j, m[e].distance_ratio, m[e].distance);
            b200m_result_free(res);
        }
        b200m_remove_view(_ctx, qId);
        return rc == B200M_OK && !vec_putative_matches.empty();
    }

  private:
    b200m_ctx* _ctx = nullptr;
    int _dtype = -1;
    uint32_t _dbId = 0;
    bool _built = false;
};

/// The cases an integration adds to createRegionsMatcher (RegionsMatcher.cpp:54-176) for BRUTE_FORCE_L2_B200 /
/// BRUTE_FORCE_HAMMING_B200, with the factory's own validity rules (:61-64): a scalar Regions with the Hamming matcher or
/// a binary Regions with an L2 matcher yields a null matcher.
inline std::unique_ptr<IRegionsMatcher> createRegionsMatcher_b200(std::mt19937& randomNumberGenerator, const feature::Regions& regions, bool hammingMatcher)
{
    std::unique_ptr<IRegionsMatcher> out;
    if (regions.IsScalar() && hammingMatcher)
        return out;
    if (regions.IsBinary() && !hammingMatcher)
        return out;
    if (b200detail::regionsDtype(regions) < 0)
        return out;
    out.reset(new RegionsMatcher_b200(randomNumberGenerator, regions));
    return out;
}

}  // namespace matching
}  // namespace aliceVision
