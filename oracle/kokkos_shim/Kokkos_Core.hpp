// Kokkos_Core.hpp -- a minimal, serial/OpenMP stand-in for the parts of Kokkos 4.7 that the
// reference's direct-sum headers use.  TEST INFRASTRUCTURE (oracle/_ref): it exists only so that
// /root/reference/src/lpm_{sphere_functions,bve_sphere_kernels,incompressible2d_kernels,swe_kernels}.hpp
// can be compiled IN PLACE (never copied) and run on the CPU to pin our restatement (oracle/lpm_oracle.c).
// This is our own code; it is not derived from Kokkos sources.  Semantics kept:
//   * View<T*...>: reference-counted shallow copies, zero-initialised, LayoutRight;
//   * parallel_for(TeamPolicy(n, AUTO), f): one team of size 1 per league rank (what Kokkos-OpenMP
//     resolves AUTO to without SMT), league ranks spread over OpenMP threads;
//   * parallel_reduce(TeamThreadRange/TeamVectorRange(member, n), functor, result): sequential
//     j = 0..n-1 into an identity-initialised value that OVERWRITES result.
#ifndef ORACLE_KOKKOS_SHIM_CORE_HPP
#define ORACLE_KOKKOS_SHIM_CORE_HPP

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <initializer_list>
#include <limits>
#include <string>
#include <type_traits>
#include <utility>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_VERSION 40700

namespace Kokkos {

struct LayoutRight {};
struct LayoutLeft {};
struct Serial {
  using execution_space = Serial;
  struct memory_space_t {};
  using memory_space = memory_space_t;
};
struct HostSpace {
  using execution_space = Serial;
  using memory_space = HostSpace;
};
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;
template <class E, class M>
struct Device {
  using execution_space = E;
  using memory_space = M;
};

template <class Space, class Mem>
struct SpaceAccessibility {  // everything lives on the host here
  enum : bool { accessible = true, assignable = true, deepcopy = true };
};

struct ALL_t {
  constexpr ALL_t operator()() const { return ALL_t{}; }
};
constexpr ALL_t ALL{};
struct AUTO_t {
  constexpr AUTO_t operator()() const { return AUTO_t{}; }
};
constexpr AUTO_t AUTO{};

inline void abort(const char* msg) {
  std::fprintf(stderr, "Kokkos::abort: %s\n", msg);
  std::abort();
}
inline void fence() {}
inline void fence(const std::string&) {}
namespace Profiling {
inline void pushRegion(const std::string&) {}
inline void popRegion() {}
}  // namespace Profiling

// ---- Array -------------------------------------------------------------------------------------
template <class T, std::size_t N>
struct Array {
  T m_internal_implementation_private_member_data[N];
  KOKKOS_INLINE_FUNCTION T& operator[](std::size_t i) { return m_internal_implementation_private_member_data[i]; }
  KOKKOS_INLINE_FUNCTION const T& operator[](std::size_t i) const {
    return m_internal_implementation_private_member_data[i];
  }
  KOKKOS_INLINE_FUNCTION T* data() { return m_internal_implementation_private_member_data; }
  KOKKOS_INLINE_FUNCTION const T* data() const { return m_internal_implementation_private_member_data; }
  static constexpr std::size_t size() { return N; }
};

template <class T>
struct reduction_identity;
template <>
struct reduction_identity<double> {
  static double sum() { return 0.0; }
  static double prod() { return 1.0; }
};
template <>
struct reduction_identity<int> {
  static int sum() { return 0; }
};

// ---- View --------------------------------------------------------------------------------------
namespace Impl {
template <class T>
struct ViewTraits {  // rank-0
  using value_type = T;
  static constexpr int rank = 0;
  static constexpr int dyn = 0;
  static std::size_t ext(int, const std::size_t*) { return 1; }
};
template <class T>
struct ViewTraits<T*> {
  using value_type = typename ViewTraits<T>::value_type;
  static constexpr int rank = ViewTraits<T>::rank + 1;
  static constexpr int dyn = ViewTraits<T>::dyn + 1;
};
template <class T, std::size_t N>
struct ViewTraits<T[N]> {
  using value_type = typename ViewTraits<T>::value_type;
  static constexpr int rank = ViewTraits<T>::rank + 1;
  static constexpr int dyn = ViewTraits<T>::dyn;
};
// static extents, outermost first
template <class T>
struct StaticExt {
  static void fill(std::size_t*, int) {}
};
template <class T>
struct StaticExt<T*> {
  static void fill(std::size_t* e, int k) {
    e[k] = 0;
    StaticExt<T>::fill(e, k + 1);
  }
};
template <class T, std::size_t N>
struct StaticExt<T[N]> {
  // C++ array-of-array types list the OUTER static extent first: (T[M])[N] is T[N][M]
  static void fill(std::size_t* e, int k) {
    e[k] = N;
    StaticExt<T>::fill(e, k + 1);
  }
};
}  // namespace Impl

template <class T>
struct Slice1;

template <class DataType, class... Props>
class View {
 public:
  using traits = Impl::ViewTraits<DataType>;
  using value_type = typename traits::value_type;
  using non_const_value_type = typename std::remove_const<value_type>::type;
  using HostMirror = View<DataType, Props...>;
  using execution_space = Serial;
  using memory_space = HostSpace;
  using host_mirror_type = HostMirror;
  static constexpr int rank = traits::rank;
  static constexpr int Rank = traits::rank;

  View() { init_static(); }
  explicit View(const std::string&, std::size_t n0 = 0, std::size_t n1 = 0, std::size_t n2 = 0) {
    init_static();
    const std::size_t dyn[3] = {n0, n1, n2};
    int d = 0;
    for (int k = 0; k < rank; ++k)
      if (ext_[k] == 0 && d < traits::dyn) ext_[k] = dyn[d++];
    allocate();
  }
  // unmanaged wrap of existing memory (used by the driver to view caller arrays)
  View(value_type* p, std::size_t n0, std::size_t n1 = 0, std::size_t n2 = 0) {
    init_static();
    const std::size_t dyn[3] = {n0, n1, n2};
    int d = 0;
    for (int k = 0; k < rank; ++k)
      if (ext_[k] == 0 && d < traits::dyn) ext_[k] = dyn[d++];
    ptr_ = const_cast<non_const_value_type*>(p);
    strides();
  }
  template <class OT, class... OP>
  View(const View<OT, OP...>& o) {  // const-conversion / same-shape shallow copy
    for (int k = 0; k < 3; ++k) {
      ext_[k] = o.extent(k) * (k < rank ? 1 : 1);
    }
    own_ = o.owner();
    ptr_ = const_cast<non_const_value_type*>(o.data());
    strides();
  }

  // rows [r.first, r.second) of another view: View(v, pair) / View(v, pair, ALL) as Kokkos' subview constructor
  template <class OT, class... OP, class I>
  View(const View<OT, OP...>& o, std::pair<I, I> r) : View(o) {
    ptr_ += (std::size_t)r.first * s0_;
    ext_[0] = (std::size_t)(r.second - r.first);
  }
  template <class OT, class... OP, class I>
  View(const View<OT, OP...>& o, std::pair<I, I> r, ALL_t) : View(o, r) {}
  // a row of a rank-2 view, as an unmanaged rank-1 view
  template <class ST>
  View(const Slice1<ST>& sl) {
    init_static();
    ext_[0] = sl.n;
    ptr_ = const_cast<non_const_value_type*>(sl.p);
    strides();
  }

  KOKKOS_INLINE_FUNCTION std::size_t extent(int k) const { return k < rank ? ext_[k] : 1; }
  KOKKOS_INLINE_FUNCTION int extent_int(int k) const { return (int)extent(k); }
  KOKKOS_INLINE_FUNCTION std::size_t size() const { return ext_[0] * ext_[1] * ext_[2]; }
  KOKKOS_INLINE_FUNCTION value_type* data() const { return ptr_; }
  const std::shared_ptr<void>& owner() const { return own_; }
  std::string label() const { return ""; }
  bool is_allocated() const { return ptr_ != nullptr; }

  KOKKOS_INLINE_FUNCTION value_type& operator()() const { return ptr_[0]; }
  template <class I>
  KOKKOS_INLINE_FUNCTION value_type& operator()(I i) const {
    return ptr_[(std::size_t)i * s0_];
  }
  template <class I, class J>
  KOKKOS_INLINE_FUNCTION value_type& operator()(I i, J j) const {
    return ptr_[(std::size_t)i * s0_ + (std::size_t)j * s1_];
  }
  template <class I, class J, class K>
  KOKKOS_INLINE_FUNCTION value_type& operator()(I i, J j, K k) const {
    return ptr_[(std::size_t)i * s0_ + (std::size_t)j * s1_ + (std::size_t)k];
  }
  template <class I>
  KOKKOS_INLINE_FUNCTION value_type& operator[](I i) const {
    return ptr_[(std::size_t)i * s0_];
  }

 private:
  void init_static() {
    ext_[0] = ext_[1] = ext_[2] = 1;
    std::size_t e[4] = {1, 1, 1, 1};
    Impl::StaticExt<DataType>::fill(e, 0);
    // StaticExt lists pointer (dynamic) dims as 0 and static dims by value, but for T*[3] the C++
    // type is (T*)[3]: the static extent comes FIRST in the type although it is the LAST view
    // dimension.  Kokkos orders dynamic dimensions first, so sort zeros to the front.
    std::size_t sorted[3];
    int n = 0;
    for (int k = 0; k < rank; ++k)
      if (e[k] == 0) sorted[n++] = 0;
    for (int k = 0; k < rank; ++k)
      if (e[k] != 0) sorted[n++] = e[k];
    for (int k = 0; k < rank; ++k) ext_[k] = sorted[k];
  }
  void strides() {
    s1_ = rank >= 3 ? ext_[2] : 1;
    s0_ = rank >= 2 ? ext_[1] * s1_ : 1;
  }
  void allocate() {
    const std::size_t n = size();
    if (n > 0) {
      auto* p = static_cast<non_const_value_type*>(std::calloc(n, sizeof(non_const_value_type)));
      own_ = std::shared_ptr<void>(p, std::free);
      ptr_ = p;
    }
    strides();
  }
  std::size_t ext_[3];
  std::size_t s0_ = 1, s1_ = 1;
  std::shared_ptr<void> own_;
  non_const_value_type* ptr_ = nullptr;
};

// 1-D strided slice returned by subview(v, i, ALL)
template <class T>
struct Slice1 {
  T* p;
  std::size_t n;
  KOKKOS_INLINE_FUNCTION T& operator[](std::size_t k) const { return p[k]; }
  KOKKOS_INLINE_FUNCTION T& operator()(std::size_t k) const { return p[k]; }
  KOKKOS_INLINE_FUNCTION std::size_t extent(int) const { return n; }
  KOKKOS_INLINE_FUNCTION T* data() const { return p; }
};
template <class D, class... P, class I>
KOKKOS_INLINE_FUNCTION Slice1<typename View<D, P...>::value_type> subview(const View<D, P...>& v, I i, ALL_t) {
  return {v.data() + (std::size_t)i * v.extent(1), v.extent(1)};
}
template <class D, class... P, class I>
KOKKOS_INLINE_FUNCTION View<D, P...> subview(const View<D, P...>& v, std::pair<I, I> r, ALL_t) {
  return View<D, P...>(v, r);
}
template <class D, class... P, class I>
KOKKOS_INLINE_FUNCTION View<D, P...> subview(const View<D, P...>& v, std::pair<I, I> r) {
  return View<D, P...>(v, r);
}

template <class V>
typename V::HostMirror create_mirror_view(const V& v) {
  return v;
}
template <class A, class B, typename std::enable_if<!std::is_arithmetic<B>::value, int>::type = 0>
void deep_copy(const A& dst, const B& src) {
  if ((const void*)dst.data() != (const void*)src.data())
    std::memcpy((void*)dst.data(), (const void*)src.data(), dst.size() * sizeof(typename A::value_type));
}
template <class A, class B, typename std::enable_if<std::is_arithmetic<B>::value, int>::type = 0>
void deep_copy(const A& dst, const B& value) {
  for (std::size_t k = 0; k < dst.size(); ++k) dst.data()[k] = (typename A::value_type)value;
}
template <class V>
typename V::HostMirror create_mirror(const V& v) {
  return v;
}

// ---- policies ----------------------------------------------------------------------------------
struct TeamMember {
  int league_rank_;
  KOKKOS_INLINE_FUNCTION int league_rank() const { return league_rank_; }
  KOKKOS_INLINE_FUNCTION int team_rank() const { return 0; }
  KOKKOS_INLINE_FUNCTION int team_size() const { return 1; }
  KOKKOS_INLINE_FUNCTION void team_barrier() const {}
};
template <class... P>
struct TeamPolicy {
  using member_type = TeamMember;
  int league_size_;
  TeamPolicy() : league_size_(0) {}
  template <class TS>
  TeamPolicy(int league, TS) : league_size_(league) {}
  template <class TS, class VL>
  TeamPolicy(int league, TS, VL) : league_size_(league) {}
};
template <class... P>
struct RangePolicy {
  long b, e;
  RangePolicy(long b_, long e_) : b(b_), e(e_) {}
};
struct TeamRange {
  long n;
};
template <class I>
KOKKOS_INLINE_FUNCTION TeamRange TeamThreadRange(const TeamMember&, I n) {
  return {(long)n};
}
template <class I>
KOKKOS_INLINE_FUNCTION TeamRange TeamVectorRange(const TeamMember&, I n) {
  return {(long)n};
}
template <class I>
KOKKOS_INLINE_FUNCTION TeamRange ThreadVectorRange(const TeamMember&, I n) {
  return {(long)n};
}

// ---- parallel_for ------------------------------------------------------------------------------
template <class F, class... P>
void parallel_for(const std::string&, const TeamPolicy<P...>& pol, const F& f) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < pol.league_size_; ++i) f(TeamMember{i});
}
template <class F, class... P>
void parallel_for(const TeamPolicy<P...>& pol, const F& f) {
  parallel_for(std::string(), pol, f);
}
template <class F, class... P>
void parallel_for(const std::string&, const RangePolicy<P...>& pol, const F& f) {
#pragma omp parallel for schedule(static)
  for (long i = pol.b; i < pol.e; ++i) f((int)i);
}
template <class F, class... P>
void parallel_for(const RangePolicy<P...>& pol, const F& f) {
  parallel_for(std::string(), pol, f);
}
template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_for(const std::string&, I n, const F& f) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)n; ++i) f((int)i);
}
template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_for(I n, const F& f) {
  parallel_for(std::string(), n, f);
}
// nested (inside a team)
template <class F>
KOKKOS_INLINE_FUNCTION void parallel_for(const TeamRange& r, const F& f) {
  for (long j = 0; j < r.n; ++j) f((int)j);
}

// ---- parallel_reduce ---------------------------------------------------------------------------
template <class V>
KOKKOS_INLINE_FUNCTION V reduce_identity_of() {
  return V();  // Tuple default-constructs to zero; double() == 0
}
// nested: sequential j, result overwritten
template <class F, class V>
KOKKOS_INLINE_FUNCTION void parallel_reduce(const TeamRange& r, const F& f, V& result) {
  V acc = reduce_identity_of<V>();
  for (long j = 0; j < r.n; ++j) f((int)j, acc);
  result = acc;
}
// top level over an integer range (host diagnostics): sequential to keep it deterministic
template <class I, class F, class V, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(const std::string&, I n, const F& f, V& result) {
  V acc = reduce_identity_of<V>();
  for (long j = 0; j < (long)n; ++j) f((int)j, acc);
  result = acc;
}
template <class I, class F, class V, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(I n, const F& f, V& result) {
  parallel_reduce(std::string(), n, f, result);
}


// reducer objects passed by value as the last argument (Kokkos::Max<Real>(result)): sequential, identity-initialised
template <class T>
struct Max {
  T& ref;
  explicit Max(T& r) : ref(r) {}
};
template <class I, class F, class T, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(I n, const F& f, Max<T> red) {
  T acc = std::numeric_limits<T>::lowest();
  for (long j = 0; j < (long)n; ++j) f((int)j, acc);
  red.ref = acc;
}

template <class I, class F, class T, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(const std::string&, I n, const F& f, Max<T> red) {
  parallel_reduce(n, f, red);
}
template <class T>
struct Min {
  T& ref;
  explicit Min(T& r) : ref(r) {}
};
template <class I, class F, class T, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(I n, const F& f, Min<T> red) {
  T acc = std::numeric_limits<T>::max();
  for (long j = 0; j < (long)n; ++j) f((int)j, acc);
  red.ref = acc;
}
template <class I, class F, class T, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(const std::string&, I n, const F& f, Min<T> red) {
  parallel_reduce(n, f, red);
}
template <class T>
struct MinMaxScalar {
  T min_val, max_val;
};
template <class T>
struct MinMax {
  using value_type = MinMaxScalar<T>;
  value_type& ref;
  explicit MinMax(value_type& r) : ref(r) {}
};
template <class I, class F, class T, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(I n, const F& f, MinMax<T> red) {
  MinMaxScalar<T> acc{std::numeric_limits<T>::max(), std::numeric_limits<T>::lowest()};
  for (long j = 0; j < (long)n; ++j) f((int)j, acc);
  red.ref = acc;
}
template <class I, class F, class T, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_reduce(const std::string&, I n, const F& f, MinMax<T> red) {
  parallel_reduce(n, f, red);
}

// MDRangePolicy<Rank<2>>({b0, b1}, {e0, e1}) with a sum reduction (host diagnostics: nan counts), row-major, sequential
template <int N>
struct Rank {
  static constexpr int rank = N;
};
template <class R>
struct MDRangePolicy {
  long b[3], e[3];
  MDRangePolicy(std::initializer_list<long> lo, std::initializer_list<long> hi) {
    int k = 0;
    for (long v : lo) b[k++] = v;
    k = 0;
    for (long v : hi) e[k++] = v;
  }
};
template <class R, class F, class V>
void parallel_reduce(const MDRangePolicy<R>& pol, const F& f, V& result) {
  V acc = V();
  for (long i = pol.b[0]; i < pol.e[0]; ++i)
    for (long j = pol.b[1]; j < pol.e[1]; ++j) f((int)i, (int)j, acc);
  result = acc;
}
template <class R, class F, class V>
void parallel_reduce(const std::string&, const MDRangePolicy<R>& pol, const F& f, V& result) {
  parallel_reduce(pol, f, result);
}
template <class R, class F>
void parallel_for(const MDRangePolicy<R>& pol, const F& f) {
  for (long i = pol.b[0]; i < pol.e[0]; ++i)
    for (long j = pol.b[1]; j < pol.e[1]; ++j) f((int)i, (int)j);
}
template <class R, class F>
void parallel_for(const std::string&, const MDRangePolicy<R>& pol, const F& f) {
  parallel_for(pol, f);
}

// exclusive/inclusive scan protocol of Kokkos: f(i, partial, is_final), sequential here
template <class I, class F, class V, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_scan(const std::string&, I n, const F& f, V& result) {
  V acc = V();
  for (long j = 0; j < (long)n; ++j) f((int)j, acc, true);
  result = acc;
}
template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_scan(const std::string&, I n, const F& f) {
  long long acc = 0;
  (void)acc;
  int a = 0;
  for (long j = 0; j < (long)n; ++j) f((int)j, a, true);
}
template <class I, class F, class V, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
void parallel_scan(I n, const F& f, V& result) {
  parallel_scan(std::string(), n, f, result);
}

inline void initialize(int&, char**) {}
inline void initialize() {}
inline void finalize() {}

}  // namespace Kokkos

#endif
