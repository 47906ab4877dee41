// When to write an image: reina::tools::SaveManager (src/tools/SaveManager.cpp:6-44) restated.
// Thresholds fire in ascending order, one per call, sample thresholds before time thresholds, and only once the
// counter is STRICTLY greater than the threshold; the file is named after the threshold, not after the actual count
// ("output_<N>spp.png", "output_<T>sec.png" with T truncated to an integer).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace rbhost {

struct SaveInfo {
    bool shouldSave;
    std::string filename;
};

class SaveManager {
public:
    SaveManager() = default;
    SaveManager(std::vector<int> samples, std::vector<double> times)
        : saveTimes(std::move(times)), saveSamples(std::move(samples)) {
        std::sort(saveSamples.begin(), saveSamples.end(), std::greater<>());
        std::sort(saveTimes.begin(), saveTimes.end(), std::greater<>());
    }

    SaveInfo shouldSave(uint32_t samples, double time) {
        // the int threshold is converted to unsigned for the comparison, as in the reference's `int < uint32_t`
        if (!saveSamples.empty() && static_cast<uint32_t>(saveSamples.back()) < samples) {
            const int s = saveSamples.back();
            saveSamples.pop_back();
            return {true, "output_" + std::to_string(s) + "spp.png"};
        }
        if (!saveTimes.empty() && saveTimes.back() < time) {
            const double t = saveTimes.back();
            saveTimes.pop_back();
            return {true, "output_" + std::to_string(static_cast<int>(t)) + "sec.png"};
        }
        return {false, ""};
    }

    bool pending() const { return !saveSamples.empty() || !saveTimes.empty(); }
    bool pendingSamples() const { return !saveSamples.empty(); }

private:
    std::vector<double> saveTimes;
    std::vector<int> saveSamples;
};

}  // namespace rbhost
