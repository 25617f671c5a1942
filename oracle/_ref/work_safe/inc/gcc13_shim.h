// Pre-included (-include) when compiling the UNMODIFIED reference sources with g++ 13.
// Test infrastructure only (oracle/_ref build); not product code.
#pragma once
#include <tuple>
#include <cstddef>
namespace PackElementDetail {
    // g++ 13 has no __type_pack_element; the reference names this fallback
    // (Core/TypePack.h:L66-75) without defining it.
    template<std::size_t I, class... Ts>
    struct TypePackElementT { using Type = std::tuple_element_t<I, std::tuple<Ts...>>; };
}
#include "Core/Types.h"
// g++ 13: no CTAD through the aggregate/inherited ctor of Pair
template<class F, class S>
Pair(F&&, S&&) -> Pair<std::remove_cvref_t<F>, std::remove_cvref_t<S>>;
