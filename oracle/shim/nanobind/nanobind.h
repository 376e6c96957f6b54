// oracle/shim/nanobind/nanobind.h -- TEST INFRASTRUCTURE ONLY.
//
// A stand-in for the slice of nanobind that mdapy's hot-path translation
// units use (src/type.h:9-21 and the NB_MODULE blocks at the end of each
// src/*.cpp).  It lets the reference sources compile UNMODIFIED, from where
// they lie under /root/reference, into oracle/_ref/libmdapy_ref.so so that
// the parity tests can call the real reference arithmetic through ctypes.
// Nothing here is shipped in the product library.
//
// Provided: nb::ndarray<...> (raw pointer + up to 3 extents, row-major),
// its .view(), nb::capsule (ownership is handed to the C wrapper, never
// freed here), nb::arg, nb::module_/class_/init (no-ops) and NB_MODULE.
#pragma once
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <type_traits>
#include <tuple>
#include <string>

namespace nanobind {

struct ro {};
struct numpy {};
template <int N> struct ndim { static constexpr int value = N; };

struct capsule {
    void *ptr{nullptr};
    void (*deleter)(void *) noexcept {nullptr};
    capsule() = default;
    capsule(void *p, void (*d)(void *) noexcept) : ptr(p), deleter(d) {}
};

struct arg {
    explicit arg(const char *) {}
};

namespace detail {
// first template argument that is not one of our tags is the scalar type
template <class T> struct is_tag : std::false_type {};
template <> struct is_tag<ro> : std::true_type {};
template <> struct is_tag<numpy> : std::true_type {};
template <int N> struct is_tag<ndim<N>> : std::true_type {};

template <class... A> struct pick_scalar;
template <class H, class... T> struct pick_scalar<H, T...> {
    using type = std::conditional_t<is_tag<H>::value, typename pick_scalar<T...>::type, H>;
};
template <> struct pick_scalar<> { using type = void; };

template <class... A> struct has_ro : std::bool_constant<(std::is_same_v<A, ro> || ...)> {};

template <class T> struct strided_view {
    T *p;
    size_t ext[3];
    inline size_t shape(size_t i) const { return ext[i]; }
    inline T &operator()(size_t i) const { return p[i]; }
    inline T &operator()(size_t i, size_t j) const { return p[i * ext[1] + j]; }
    inline T &operator()(size_t i, size_t j, size_t k) const { return p[(i * ext[1] + j) * ext[2] + k]; }
    inline T *data() const { return p; }
};
} // namespace detail

template <class... Args> class ndarray {
    using raw_scalar = typename detail::pick_scalar<Args...>::type;

public:
    using Scalar = std::conditional_t<detail::has_ro<Args...>::value, const raw_scalar, raw_scalar>;

    ndarray() = default;
    ndarray(Scalar *p, std::initializer_list<size_t> shp, capsule own = {}) : ptr_(p), owner_(own) {
        nd_ = 0;
        for (size_t s : shp) ext_[nd_++] = s;
        for (int i = nd_; i < 3; ++i) ext_[i] = 1;
    }
    // mutable -> read-only conversion, as nanobind allows
    template <class... B, class = std::enable_if_t<std::is_convertible_v<typename ndarray<B...>::Scalar *, Scalar *>>>
    ndarray(const ndarray<B...> &o) : ptr_(o.data()), nd_(o.ndim_()) {
        for (int i = 0; i < 3; ++i) ext_[i] = o.shape(i);
    }

    inline size_t shape(size_t i) const { return ext_[i]; }
    inline size_t size() const {
        size_t s = 1;
        for (int i = 0; i < nd_; ++i) s *= ext_[i];
        return nd_ ? s : 0;
    }
    inline Scalar *data() const { return ptr_; }
    inline int ndim_() const { return nd_; }
    inline Scalar &operator()(size_t i) const { return ptr_[i]; }
    inline Scalar &operator()(size_t i, size_t j) const { return ptr_[i * ext_[1] + j]; }
    inline Scalar &operator()(size_t i, size_t j, size_t k) const { return ptr_[(i * ext_[1] + j) * ext_[2] + k]; }
    inline detail::strided_view<Scalar> view() const { return {ptr_, {ext_[0], ext_[1], ext_[2]}}; }
    const capsule &owner() const { return owner_; }

private:
    Scalar *ptr_{nullptr};
    size_t ext_[3]{0, 1, 1};
    int nd_{0};
    capsule owner_{};
};

template <class... A> struct init {};

template <class C> struct class_ {
    template <class M> class_(M &, const char *) {}
    template <class... X> class_ &def(X &&...) { return *this; }
};

struct module_ {
    template <class... X> module_ &def(X &&...) { return *this; }
    const char *&doc() {
        static const char *d = nullptr;
        return d;
    }
};

} // namespace nanobind

#define NB_MODULE(name, var) [[maybe_unused]] static void nb_shim_init_##name(nanobind::module_ &var)
