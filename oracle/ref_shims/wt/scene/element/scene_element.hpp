// ORACLE -- TEST INFRASTRUCTURE ONLY.  Build shim for oracle/_ref: the scene-element base class (include/wt/scene/element/scene_element.hpp: id,
// description for the GUI, loader hooks) reduced to what sampler.hpp and fsd_sampler.hpp derive from.
#pragma once
#include <string>
namespace wt::scene {
namespace element { struct info_t { std::string id, type; }; }
namespace loader { class loader_t; struct node_t; }
class scene_element_t {
    std::string id_;
public:
    explicit scene_element_t(std::string id) : id_(std::move(id)) {}
    scene_element_t(scene_element_t&&) = default;
    scene_element_t(const scene_element_t&) = default;
    virtual ~scene_element_t() noexcept = default;
    [[nodiscard]] const std::string& get_id() const noexcept { return id_; }
    [[nodiscard]] virtual element::info_t description() const = 0;
};
}
