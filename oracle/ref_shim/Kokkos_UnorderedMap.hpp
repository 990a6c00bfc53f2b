/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for Kokkos::UnorderedMap: the
 * insert / find / exists / value_at subset the reference's HypreLinearSystem
 * uses; copies share their state (as Kokkos' reference-counted maps do). */
#ifndef NW_REF_SHIM_KOKKOS_UNORDEREDMAP_HPP
#define NW_REF_SHIM_KOKKOS_UNORDEREDMAP_HPP
#include <Kokkos_Core.hpp>
#include <cstdint>
#include <unordered_map>
namespace Kokkos {
struct UnorderedMapInsertResult
{
  bool ok, had;
  uint32_t idx;
  bool success() const { return ok; }
  bool existing() const { return had; }
  bool failed() const { return !ok && !had; }
  uint32_t index() const { return idx; }
};
template <class Key, class Value, class... P>
class UnorderedMap
{
  struct State
  {
    std::unordered_map<Key, uint32_t> index;
    std::vector<Key> keys;
    std::vector<Value> values;
  };
  std::shared_ptr<State> st_;

public:
  using HostMirror = UnorderedMap;
  using size_type = uint32_t;
  static constexpr uint32_t invalid_index = ~0u;
  UnorderedMap(size_t = 0) : st_(std::make_shared<State>()) {}
  UnorderedMapInsertResult insert(const Key& k, const Value& v = Value()) const
  {
    auto it = st_->index.find(k);
    if (it != st_->index.end())
      return UnorderedMapInsertResult{false, true, it->second};
    const uint32_t i = (uint32_t)st_->keys.size();
    st_->index.emplace(k, i);
    st_->keys.push_back(k);
    st_->values.push_back(v);
    return UnorderedMapInsertResult{true, false, i};
  }
  uint32_t find(const Key& k) const
  {
    auto it = st_->index.find(k);
    return it == st_->index.end() ? invalid_index : it->second;
  }
  bool exists(const Key& k) const { return st_->index.count(k) != 0; }
  bool valid_at(uint32_t i) const { return i < st_->keys.size(); }
  Value& value_at(uint32_t i) const { return st_->values[i]; }
  const Key& key_at(uint32_t i) const { return st_->keys[i]; }
  uint32_t size() const { return (uint32_t)st_->keys.size(); }
  uint32_t capacity() const { return (uint32_t)st_->keys.size() + 1024u; }
  bool rehash(size_t = 0) { return true; }
  void clear() { *st_ = State(); }
  bool failed_insert() const { return false; }
  void copy_from(const UnorderedMap& o) { *st_ = *o.st_; }
};
template <class K, class V, class... P1, class... P2>
void
deep_copy(UnorderedMap<K, V, P1...>& dst, const UnorderedMap<K, V, P2...>& src)
{
  dst.copy_from(src);
}
} // namespace Kokkos
#endif
