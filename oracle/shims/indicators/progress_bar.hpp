// Shim for indicators 2.3 (cosmetic progress bar; src/simulation.cpp:703-709,737-738,744). No-op.
#pragma once
#include <cstddef>
#include <string>
#include <vector>
namespace indicators {
enum class Color { cyan };
enum class FontStyle { bold };
namespace option {
struct BarWidth { std::size_t v; };
struct Start { std::string v; };
struct Fill { std::string v; };
struct Lead { std::string v; };
struct Remainder { std::string v; };
struct End { std::string v; };
struct PrefixText { std::string v; };
struct PostfixText { std::string v; };
struct ForegroundColor { Color v; };
struct ShowElapsedTime { bool v; };
struct ShowRemainingTime { bool v; };
struct FontStyles { std::vector<FontStyle> v; };
struct MaxProgress { std::size_t v; };
}  // namespace option
class ProgressBar {
public:
    template <typename... A> explicit ProgressBar(A &&...) {}
    template <typename O> void set_option(O &&) {}
    void tick() {}
};
}  // namespace indicators
