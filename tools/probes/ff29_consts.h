// Generated from the BN254 moduli (fawkes-crypto/src/engines/bn256/mod.rs:13,23); 9 x 29-bit limbs.
#pragma once
#include <cstdint>
namespace fb {
namespace c29 {
// FQ: p, 2p, -p^-1 mod 2^29, ONE = 2^261 mod p, TO_STD = 2^256 mod p (plain), TO_INT32 = 2^261 mod p in 2^256-Montgomery form (8 x u32)
#define FB_C29_FQ_P {0x187cfd47u, 0x010460b6u, 0x1c72a34fu, 0x02d522d0u, 0x1585d978u, 0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu}
#define FB_C29_FQ_2P {0x10f9fa8eu, 0x0208c16du, 0x18e5469eu, 0x05aa45a1u, 0x0b0bb2f0u, 0x05b68181u, 0x014dc282u, 0x1cb84c68u, 0x0060c89cu}
#define FB_C29_FQ_PINV 0x04866389u
#define FB_C29_FQ_ONE {0x157ccc21u, 0x141c2758u, 0x185230d3u, 0x014c0419u, 0x0aa36fb9u, 0x1d4240ceu, 0x11d54c07u, 0x052ac7a8u, 0x000dc836u}
#define FB_C29_FQ_TO_STD {0x058f0d9du, 0x1aea1c6eu, 0x11c2cf74u, 0x11d651ebu, 0x1462c0a7u, 0x11b7bc3cu, 0x1cbd99bau, 0x183340fbu, 0x000e0a77u}
#define FB_C29_FQ_TO_INT_W {0x157ccc21u, 0x4e8384ebu, 0x0ce148c3u, 0xfb90a602u, 0x819caa36u, 0x5301fa84u, 0x563d4475u, 0x0dc83629u}
// FR: p, 2p, -p^-1 mod 2^29, ONE = 2^261 mod p, TO_STD = 2^256 mod p (plain), TO_INT32 = 2^261 mod p in 2^256-Montgomery form (8 x u32)
#define FB_C29_FR_P {0x10000001u, 0x1f0fac9fu, 0x0e5c2450u, 0x07d090f3u, 0x1585d283u, 0x02db40c0u, 0x00a6e141u, 0x0e5c2634u, 0x0030644eu}
#define FB_C29_FR_2P {0x00000002u, 0x1e1f593fu, 0x1cb848a1u, 0x0fa121e6u, 0x0b0ba506u, 0x05b68181u, 0x014dc282u, 0x1cb84c68u, 0x0060c89cu}
#define FB_C29_FR_PINV 0x0fffffffu
#define FB_C29_FR_ONE {0x0fffff57u, 0x1ea70ab4u, 0x052c068bu, 0x17504f49u, 0x0aa8075bu, 0x1d4240ceu, 0x11d54c07u, 0x052ac7a8u, 0x000dc836u}
#define FB_C29_FR_TO_STD {0x0ffffffbu, 0x04b1a0e2u, 0x18334a6bu, 0x18ed2b3eu, 0x1462e36fu, 0x11b7bc3cu, 0x1cbd99bau, 0x183340fbu, 0x000e0a77u}
#define FB_C29_FR_TO_INT_W {0x8fffff57u, 0x2fd4e156u, 0xa494b01au, 0x75bba827u, 0x819caa80u, 0x5301fa84u, 0x563d4475u, 0x0dc83629u}
}  // namespace c29
}  // namespace fb
