/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <Kokkos_Core.hpp>: an
 * unmanaged / shared-pointer-managed row-major View of rank <= 8, Array, the
 * space tags and abort -- what the reference's master-element sources touch.
 * Serial host semantics only.  Not Kokkos code; written for this repo. */
#ifndef NW_REF_SHIM_KOKKOS_CORE_HPP
#define NW_REF_SHIM_KOKKOS_CORE_HPP
#include "Kokkos_Macros.hpp"
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

namespace Kokkos {

struct LayoutRight {};
struct LayoutLeft {};
struct LayoutStride {};
enum MemoryTraitsFlags : unsigned {
  Unmanaged = 0x01,
  RandomAccess = 0x02,
  Atomic = 0x04,
  Restrict = 0x08,
  Aligned = 0x10
};
template <unsigned F>
struct MemoryTraits {};
using MemoryUnmanaged = MemoryTraits<Unmanaged>;

struct HostSpace {};
template <class Exec>
struct ScratchMemorySpace {};
struct Serial
{
  using execution_space = Serial;
  using memory_space = HostSpace;
  using scratch_memory_space = ScratchMemorySpace<Serial>;
};
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;

[[noreturn]] inline void
abort(const char* msg)
{
  std::fprintf(stderr, "Kokkos::abort: %s\n", msg);
  std::abort();
}

template <class T, size_t N>
struct Array
{
  T m_data[N ? N : 1];
  KOKKOS_INLINE_FUNCTION T& operator[](size_t i) { return m_data[i]; }
  KOKKOS_INLINE_FUNCTION const T& operator[](size_t i) const { return m_data[i]; }
  static constexpr size_t size() { return N; }
  T* data() { return m_data; }
  const T* data() const { return m_data; }
};

namespace shim {
template <class T>
struct data_type
{
  static constexpr int rank = 0;
  using value_type = T;
  static void static_extents(size_t*, int) {}
};
template <class T>
struct data_type<T*>
{
  static constexpr int rank = 1 + data_type<T>::rank;
  using value_type = typename data_type<T>::value_type;
  static void static_extents(size_t* e, int at) { data_type<T>::static_extents(e, at + 1); }
};
template <class T, size_t N>
struct data_type<T[N]>
{
  static constexpr int rank = 1 + data_type<T>::rank;
  using value_type = typename data_type<T>::value_type;
  static void static_extents(size_t* e, int at)
  {
    e[at] = N;
    data_type<T>::static_extents(e, at + 1);
  }
};
} // namespace shim

namespace shim {
template <class... P>
struct has_layout_left : std::false_type {};
template <class P0, class... P>
struct has_layout_left<P0, P...>
  : std::conditional_t<std::is_same<P0, LayoutLeft>::value, std::true_type, has_layout_left<P...>> {};
} // namespace shim

/* row-major view, column-major when LayoutLeft is among the properties; the
 * rest of the property pack is accepted and ignored */
template <class DataType, class... Props>
class View
{
public:
  using HostMirror = View;
  using memory_space = HostSpace;
  using execution_space = Serial;
  using size_type = size_t;
  static constexpr bool layout_left = shim::has_layout_left<Props...>::value;
  std::string label() const { return "view"; }
  bool is_allocated() const { return ptr_ != nullptr; }
  using value_type = typename shim::data_type<DataType>::value_type;
  using non_const_value_type = std::remove_const_t<value_type>;
  using pointer_type = value_type*;
  static constexpr int rank = shim::data_type<DataType>::rank;
  static constexpr int Rank = rank;

  View() { init(nullptr, 0, 0, 0, 0, 0, 0, 0, 0); }
  View(
    value_type* p, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0,
    size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0)
  {
    init(p, n0, n1, n2, n3, n4, n5, n6, n7);
  }
  explicit View(
    const std::string&, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0,
    size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0)
  {
    init(nullptr, n0, n1, n2, n3, n4, n5, n6, n7);
    owned_ = std::make_shared<std::vector<non_const_value_type>>(size());
    ptr_ = owned_->data();
  }
  /* non-const -> const, any property pack */
  template <class DT2, class... P2>
  View(const View<DT2, P2...>& o)
  {
    static_assert(shim::data_type<DT2>::rank == rank, "rank mismatch");
    ptr_ = o.data();
    for (int i = 0; i < 8; ++i)
      ext_[i] = o.extent(i);
    owned_ = o.owner();
  }

  template <class... I>
  KOKKOS_INLINE_FUNCTION value_type& operator()(I... idx) const
  {
    static_assert(sizeof...(I) == rank, "index count != rank");
    const size_t ii[] = {static_cast<size_t>(idx)..., 0};
    size_t off = 0;
    if (layout_left) {
      for (int d = rank - 1; d >= 0; --d)
        off = off * ext_[d] + ii[d];
    } else {
      for (int d = 0; d < rank; ++d)
        off = off * ext_[d] + ii[d];
    }
    return ptr_[off];
  }
  KOKKOS_INLINE_FUNCTION value_type& operator[](size_t i) const { return ptr_[i]; }
  size_t extent(int i) const { return i < rank ? ext_[i] : 1; }
  int extent_int(int i) const { return (int)extent(i); }
  size_t size() const
  {
    size_t s = 1;
    for (int d = 0; d < rank; ++d)
      s *= ext_[d];
    return s;
  }
  size_t span() const { return size(); }
  value_type* data() const { return ptr_; }
  std::shared_ptr<std::vector<non_const_value_type>> owner() const { return owned_; }

private:
  void init(
    value_type* p, size_t n0, size_t n1, size_t n2, size_t n3, size_t n4,
    size_t n5, size_t n6, size_t n7)
  {
    ptr_ = p;
    const size_t dyn[8] = {n0, n1, n2, n3, n4, n5, n6, n7};
    size_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    shim::data_type<DataType>::static_extents(st, 0);
    int k = 0;
    for (int d = 0; d < 8; ++d)
      ext_[d] = d < rank ? (st[d] ? st[d] : dyn[k++]) : 1;
  }
  value_type* ptr_ = nullptr;
  size_t ext_[8];
  std::shared_ptr<std::vector<non_const_value_type>> owned_;
};


/* ---- names the reference's KokkosInterface.h mentions (serial stand-ins) ---- */
struct Dynamic {};
struct Static {};
template <class S>
struct Schedule {};
template <unsigned A, unsigned B>
struct LaunchBounds {};
struct AUTO_t {};
constexpr AUTO_t AUTO{};
struct PerTeamValue { size_t v; };
struct PerThreadValue { size_t v; };
inline PerTeamValue PerTeam(size_t v) { return PerTeamValue{v}; }
inline PerThreadValue PerThread(size_t v) { return PerThreadValue{v}; }
struct ALL_t {};
inline ALL_t ALL() { return ALL_t{}; }

struct TeamMember
{
  int league_rank() const { return 0; }
  int league_size() const { return 1; }
  int team_rank() const { return 0; }
  int team_size() const { return 1; }
  void* team_scratch(int) const { return nullptr; }
  void* thread_scratch(int) const { return nullptr; }
  void team_barrier() const {}
};
template <class... P>
struct TeamPolicy
{
  using member_type = TeamMember;
  TeamPolicy(size_t, AUTO_t) {}
  TeamPolicy(size_t, int) {}
  TeamPolicy& set_scratch_size(int, PerTeamValue, PerThreadValue = PerThreadValue{0}) { return *this; }
};
template <class... P>
struct RangePolicy
{
  size_t b, e;
  RangePolicy(size_t b_, size_t e_) : b(b_), e(e_) {}
};
/* never instantiated by the files compiled here */
template <class V, class... A>
V subview(const V& v, A...);

template <class... P, class F>
void
parallel_for(const std::string&, const RangePolicy<P...>& r, const F& f)
{
  for (size_t i = r.b; i < r.e; ++i)
    f(i);
}
template <class... P, class F>
void
parallel_for(const RangePolicy<P...>& r, const F& f)
{
  for (size_t i = r.b; i < r.e; ++i)
    f(i);
}
template <class F>
void
parallel_for(const std::string&, size_t n, const F& f)
{
  for (size_t i = 0; i < n; ++i)
    f(i);
}
template <class... P, class F, class R>
void
parallel_reduce(const std::string&, const RangePolicy<P...>& r, const F& f, R& red)
{
  for (size_t i = r.b; i < r.e; ++i)
    f(i, red);
}
template <class Space = HostSpace>
void*
kokkos_malloc(const std::string&, size_t n)
{
  return std::malloc(n);
}
template <class Space = HostSpace>
void
kokkos_free(void* p)
{
  std::free(p);
}
inline void fence() {}

template <class V>
void
deep_copy(const V& v, const typename V::non_const_value_type& x)
{
  for (size_t i = 0; i < v.size(); ++i)
    v.data()[i] = x;
}
template <class D1, class... P1, class D2, class... P2>
void
deep_copy(const View<D1, P1...>& dst, const View<D2, P2...>& src)
{
  if ((const void*)dst.data() == (const void*)src.data())
    return;
  for (size_t i = 0; i < dst.size() && i < src.size(); ++i)
    dst.data()[i] = src.data()[i];
}
/* host build: a mirror view is the view itself */
template <class D, class... P>
View<D, P...>
create_mirror_view(const View<D, P...>& v)
{
  return v;
}
template <class Space, class D, class... P>
View<D, P...>
create_mirror_view(const Space&, const View<D, P...>& v)
{
  return v;
}
template <class T>
inline void
atomic_add(T* p, const T& v)
{
  *p += v;
}
template <class D, class... P>
void
resize(View<D, P...>& v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0)
{
  View<D, P...> w("resized", n0, n1, n2);
  deep_copy(w, v);
  v = w;
}
struct WithoutInitializing_t {};
constexpr WithoutInitializing_t WithoutInitializing{};
inline std::string
view_alloc(WithoutInitializing_t, const std::string& s)
{
  return s;
}
inline std::string
view_alloc(const std::string& s, WithoutInitializing_t)
{
  return s;
}

} // namespace Kokkos
#endif
