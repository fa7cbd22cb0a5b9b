// ceres/ordered_groups.h -- shim of the reference's modified CERES/include/ceres/ordered_groups.h
// (:55-193; modification M6 of SURVEY.md 2.2: members public, element_to_group_ an unordered_map).
#ifndef SWGN_CERES_ORDERED_GROUPS_H_
#define SWGN_CERES_ORDERED_GROUPS_H_
#include <map>
#include <set>
#include <unordered_map>
#include <vector>
namespace ceres {
template <typename T>
class OrderedGroups {
 public:
  bool AddElementToGroup(const T element, const int group) {
    if (group < 0) return false;
    auto it = element_to_group_.find(element);
    if (it != element_to_group_.end()) {
      if (it->second == group) return true;
      group_to_elements_[it->second].erase(element);
      if (group_to_elements_[it->second].empty()) group_to_elements_.erase(it->second);
    }
    element_to_group_[element] = group;
    group_to_elements_[group].insert(element);
    return true;
  }
  void Clear() {
    group_to_elements_.clear();
    element_to_group_.clear();
  }
  bool Remove(const T element) {
    const int current_group = GroupId(element);
    if (current_group < 0) return false;
    group_to_elements_[current_group].erase(element);
    if (group_to_elements_[current_group].empty()) group_to_elements_.erase(current_group);
    element_to_group_.erase(element);
    return true;
  }
  int GroupId(const T element) const {
    auto it = element_to_group_.find(element);
    return it == element_to_group_.end() ? -1 : it->second;
  }
  bool IsMember(const T element) const { return element_to_group_.count(element) > 0; }
  int GroupSize(const int group) const {
    auto it = group_to_elements_.find(group);
    return it == group_to_elements_.end() ? 0 : (int)it->second.size();
  }
  int NumElements() const { return (int)element_to_group_.size(); }
  int NumGroups() const { return (int)group_to_elements_.size(); }
  int MinNonZeroGroup() const { return group_to_elements_.empty() ? -1 : group_to_elements_.begin()->first; }
  const std::map<int, std::set<T>>& group_to_elements() const { return group_to_elements_; }
  const std::unordered_map<T, int>& element_to_group() const { return element_to_group_; }
  // public in the reference's copy (M6)
  std::map<int, std::set<T>> group_to_elements_;
  std::unordered_map<T, int> element_to_group_;
};
typedef OrderedGroups<double*> ParameterBlockOrdering;
}  // namespace ceres
#endif
