#include <stdint.h>
#include <cuda_runtime.h>
#include "/root/repo/circom_cvm_b200/csrc/fr.cuh"
using fr::Fr;
__device__ __forceinline__ Fr ldw(const uint4 *wb, uint32_t row, uint64_t bs) { const uint4 lo = __ldg(wb + (uint64_t)(2 * row) * bs), hi = __ldg(wb + (uint64_t)(2 * row + 1) * bs); Fr r; r.v[0]=lo.x; r.v[1]=lo.y; r.v[2]=lo.z; r.v[3]=lo.w; r.v[4]=hi.x; r.v[5]=hi.y; r.v[6]=hi.z; r.v[7]=hi.w; return r; }
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_0(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 0
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 0u;
  }
  { // constraint 1
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 4, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 7, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 8, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 9, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 1u;
  }
  { // constraint 2
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 5, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 7, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 8, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 9, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 2u;
  }
  { // constraint 3
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 6, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 7, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 8, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 9, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 3u;
  }
  { // constraint 4
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 10, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 13, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 14, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 15, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 4u;
  }
  { // constraint 5
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 11, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 13, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 14, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 15, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 5u;
  }
  { // constraint 6
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 12, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 13, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 14, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 15, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 6u;
  }
  { // constraint 7
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 16, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 19, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 20, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 21, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 7u;
  }
  { // constraint 8
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 17, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 19, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 20, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 21, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 8u;
  }
  { // constraint 9
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 18, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 19, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 20, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 21, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 9u;
  }
  { // constraint 10
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 22, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 25, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 26, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 27, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 10u;
  }
  { // constraint 11
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 23, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 25, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 26, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 27, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 11u;
  }
  { // constraint 12
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 24, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 25, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 26, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 27, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 12u;
  }
  { // constraint 13
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 28, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 31, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 32, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 33, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 13u;
  }
  { // constraint 14
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 29, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 31, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 32, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 33, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 14u;
  }
  { // constraint 15
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 30, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 31, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 32, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 33, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 15u;
  }
  { // constraint 16
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 34, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 37, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 38, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 39, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 16u;
  }
  { // constraint 17
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 35, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 37, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 38, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 39, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 17u;
  }
  { // constraint 18
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 36, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 37, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 38, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 39, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 18u;
  }
  { // constraint 19
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 40, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 43, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 44, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 45, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 19u;
  }
  { // constraint 20
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 41, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 43, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 44, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 45, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 20u;
  }
  { // constraint 21
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 42, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 43, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 44, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 45, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 21u;
  }
  { // constraint 22
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 46, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 49, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 50, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 51, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 22u;
  }
  { // constraint 23
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 47, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 49, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 50, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 51, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 23u;
  }
  { // constraint 24
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 48, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 49, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 50, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 51, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 24u;
  }
  { // constraint 25
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 52, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 55, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 56, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 57, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 25u;
  }
  { // constraint 26
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 53, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 55, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 56, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 57, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 26u;
  }
  { // constraint 27
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 54, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 55, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 56, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 57, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 27u;
  }
  { // constraint 28
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 58, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 61, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 62, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 63, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 28u;
  }
  { // constraint 29
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 59, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 61, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 62, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 63, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 29u;
  }
  { // constraint 30
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 60, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 61, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 62, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 63, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 30u;
  }
  { // constraint 31
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 64, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 67, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 68, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 69, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 31u;
  }
  { // constraint 32
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 65, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 67, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 68, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 69, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 32u;
  }
  { // constraint 33
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 66, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 67, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 68, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 69, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 33u;
  }
  { // constraint 34
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 70, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 73, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 74, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 75, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 34u;
  }
  { // constraint 35
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 71, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 73, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 74, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 75, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 35u;
  }
  { // constraint 36
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 72, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 73, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 74, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 75, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 36u;
  }
  { // constraint 37
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 76, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 79, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 80, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 81, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 37u;
  }
  { // constraint 38
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 77, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 79, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 80, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 81, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 38u;
  }
  { // constraint 39
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 78, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 79, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 80, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 81, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 39u;
  }
  { // constraint 40
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 82, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 85, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 86, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 87, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 40u;
  }
  { // constraint 41
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 83, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 85, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 86, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 87, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 41u;
  }
  { // constraint 42
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 84, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 85, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 86, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 87, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 42u;
  }
  { // constraint 43
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 88, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 91, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 92, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 93, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 43u;
  }
  { // constraint 44
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 89, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 91, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 92, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 93, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 44u;
  }
  { // constraint 45
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 90, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 91, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 92, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 93, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 45u;
  }
  { // constraint 46
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 94, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 97, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 98, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 99, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 46u;
  }
  { // constraint 47
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 95, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 97, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 98, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 99, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 47u;
  }
  { // constraint 48
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 96, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 97, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 98, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 99, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 48u;
  }
  { // constraint 49
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 100, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 103, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 104, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 105, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 49u;
  }
  { // constraint 50
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 101, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 103, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 104, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 105, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 50u;
  }
  { // constraint 51
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 102, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 103, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 104, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 105, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 51u;
  }
  { // constraint 52
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 106, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 109, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 110, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 111, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 52u;
  }
  { // constraint 53
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 107, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 109, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 110, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 111, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 53u;
  }
  { // constraint 54
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 108, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 109, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 110, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 111, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 54u;
  }
  { // constraint 55
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 112, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 115, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 116, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 117, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 55u;
  }
  { // constraint 56
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 113, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 115, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 116, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 117, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 56u;
  }
  { // constraint 57
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 114, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 115, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 116, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 117, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 57u;
  }
  { // constraint 58
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 118, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 121, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 122, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 123, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 58u;
  }
  { // constraint 59
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 119, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 121, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 122, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 123, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 59u;
  }
  { // constraint 60
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 120, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 121, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 122, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 123, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 60u;
  }
  { // constraint 61
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 124, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 127, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 128, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 129, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 61u;
  }
  { // constraint 62
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 125, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 127, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 128, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 129, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 62u;
  }
  { // constraint 63
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 126, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 127, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 128, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 129, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 63u;
  }
  { // constraint 64
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 130, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 133, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 134, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 135, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 64u;
  }
  { // constraint 65
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 131, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 133, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 134, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 135, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 65u;
  }
  { // constraint 66
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 132, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 133, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 134, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 135, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 66u;
  }
  { // constraint 67
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 136, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 139, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 140, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 141, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 67u;
  }
  { // constraint 68
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 137, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 139, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 140, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 141, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 68u;
  }
  { // constraint 69
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 138, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 139, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 140, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 141, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 69u;
  }
  { // constraint 70
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 142, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 145, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 146, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 147, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 70u;
  }
  { // constraint 71
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 143, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 145, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 146, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 147, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 71u;
  }
  { // constraint 72
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 144, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 145, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 146, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 147, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 72u;
  }
  { // constraint 73
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 148, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 151, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 152, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 153, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 73u;
  }
  { // constraint 74
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 149, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 151, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 152, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 153, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 74u;
  }
  { // constraint 75
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 150, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 151, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 152, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 153, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 75u;
  }
  { // constraint 76
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 154, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 157, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 158, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 159, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 76u;
  }
  { // constraint 77
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 155, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 157, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 158, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 159, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 77u;
  }
  { // constraint 78
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 156, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 157, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 158, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 159, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 78u;
  }
  { // constraint 79
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 160, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 163, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 164, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 165, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 79u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_1(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 80
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 161, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 163, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 164, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 165, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 80u;
  }
  { // constraint 81
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 162, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 163, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 164, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 165, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 81u;
  }
  { // constraint 82
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 166, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 169, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 170, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 171, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 82u;
  }
  { // constraint 83
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 167, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 169, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 170, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 171, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 83u;
  }
  { // constraint 84
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 168, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 169, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 170, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 171, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 84u;
  }
  { // constraint 85
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 172, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 175, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 176, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 177, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 85u;
  }
  { // constraint 86
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 173, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 175, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 176, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 177, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 86u;
  }
  { // constraint 87
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 174, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 175, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 176, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 177, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 87u;
  }
  { // constraint 88
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 178, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 181, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 182, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 183, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 88u;
  }
  { // constraint 89
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 179, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 181, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 182, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 183, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 89u;
  }
  { // constraint 90
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 180, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 181, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 182, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 183, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 90u;
  }
  { // constraint 91
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 184, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 187, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 188, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 189, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 91u;
  }
  { // constraint 92
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 185, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 187, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 188, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 189, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 92u;
  }
  { // constraint 93
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 186, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 187, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 188, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 189, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 93u;
  }
  { // constraint 94
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 190, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 193, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 194, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 195, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 94u;
  }
  { // constraint 95
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 191, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 193, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 194, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 195, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 95u;
  }
  { // constraint 96
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 192, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 193, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 194, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 195, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 96u;
  }
  { // constraint 97
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 196, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 199, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 200, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 201, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 97u;
  }
  { // constraint 98
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 197, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 199, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 200, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 201, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 98u;
  }
  { // constraint 99
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 198, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 199, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 200, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 201, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 99u;
  }
  { // constraint 100
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 202, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 205, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 206, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 207, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 100u;
  }
  { // constraint 101
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 203, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 205, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 206, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 207, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 101u;
  }
  { // constraint 102
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 204, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 205, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 206, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 207, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 102u;
  }
  { // constraint 103
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 208, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 211, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 212, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 213, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 103u;
  }
  { // constraint 104
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 209, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 211, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 212, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 213, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 104u;
  }
  { // constraint 105
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 210, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 211, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 212, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 213, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 105u;
  }
  { // constraint 106
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 214, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 217, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 218, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 219, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 106u;
  }
  { // constraint 107
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 215, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 217, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 218, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 219, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 107u;
  }
  { // constraint 108
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 216, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 217, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 218, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 219, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 108u;
  }
  { // constraint 109
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 220, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 223, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 224, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 225, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 109u;
  }
  { // constraint 110
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 221, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 223, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 224, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 225, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 110u;
  }
  { // constraint 111
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 222, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 223, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 224, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 225, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 111u;
  }
  { // constraint 112
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 226, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 229, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 230, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 231, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 112u;
  }
  { // constraint 113
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 227, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 229, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 230, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 231, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 113u;
  }
  { // constraint 114
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 228, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 229, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 230, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 231, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 114u;
  }
  { // constraint 115
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 232, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 235, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 236, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 237, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 115u;
  }
  { // constraint 116
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 233, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 235, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 236, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 237, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 116u;
  }
  { // constraint 117
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 234, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 235, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 236, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 237, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 117u;
  }
  { // constraint 118
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 238, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 241, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 242, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 243, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 118u;
  }
  { // constraint 119
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 239, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 241, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 242, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 243, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 119u;
  }
  { // constraint 120
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 240, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 241, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 242, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 243, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 120u;
  }
  { // constraint 121
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 244, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 247, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 248, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 249, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 121u;
  }
  { // constraint 122
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 245, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 247, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 248, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 249, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 122u;
  }
  { // constraint 123
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 246, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 247, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 248, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 249, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 123u;
  }
  { // constraint 124
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 250, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 253, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 254, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 255, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 124u;
  }
  { // constraint 125
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 251, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 253, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 254, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 255, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 125u;
  }
  { // constraint 126
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 252, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 253, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 254, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 255, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 126u;
  }
  { // constraint 127
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 256, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 259, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 260, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 261, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 127u;
  }
  { // constraint 128
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 257, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 259, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 260, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 261, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 128u;
  }
  { // constraint 129
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 258, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 259, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 260, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 261, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 129u;
  }
  { // constraint 130
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 262, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 265, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 266, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 267, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 130u;
  }
  { // constraint 131
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 263, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 265, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 266, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 267, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 131u;
  }
  { // constraint 132
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 264, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 265, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 266, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 267, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 132u;
  }
  { // constraint 133
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 268, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 271, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 272, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 273, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 133u;
  }
  { // constraint 134
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 269, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 271, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 272, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 273, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 134u;
  }
  { // constraint 135
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 270, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 271, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 272, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 273, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 135u;
  }
  { // constraint 136
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 274, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 277, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 278, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 279, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 136u;
  }
  { // constraint 137
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 275, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 277, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 278, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 279, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 137u;
  }
  { // constraint 138
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 276, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 277, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 278, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 279, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 138u;
  }
  { // constraint 139
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 280, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 283, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 284, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 285, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 139u;
  }
  { // constraint 140
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 281, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 283, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 284, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 285, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 140u;
  }
  { // constraint 141
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 282, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 283, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 284, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 285, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 141u;
  }
  { // constraint 142
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 286, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 289, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 290, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 291, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 142u;
  }
  { // constraint 143
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 287, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 289, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 290, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 291, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 143u;
  }
  { // constraint 144
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 288, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 289, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 290, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 291, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 144u;
  }
  { // constraint 145
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 292, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 295, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 296, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 297, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 145u;
  }
  { // constraint 146
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 293, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 295, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 296, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 297, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 146u;
  }
  { // constraint 147
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 294, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 295, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 296, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 297, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 147u;
  }
  { // constraint 148
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 298, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 301, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 302, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 303, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 148u;
  }
  { // constraint 149
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 299, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 301, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 302, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 303, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 149u;
  }
  { // constraint 150
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 300, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 301, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 302, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 303, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 150u;
  }
  { // constraint 151
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 304, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 307, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 308, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 309, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 151u;
  }
  { // constraint 152
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 305, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 307, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 308, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 309, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 152u;
  }
  { // constraint 153
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 306, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 307, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 308, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 309, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 153u;
  }
  { // constraint 154
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 310, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 313, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 314, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 315, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 154u;
  }
  { // constraint 155
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 311, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 313, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 314, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 315, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 155u;
  }
  { // constraint 156
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 312, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 313, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 314, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 315, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 156u;
  }
  { // constraint 157
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 316, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 319, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 320, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 321, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 157u;
  }
  { // constraint 158
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 317, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 319, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 320, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 321, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 158u;
  }
  { // constraint 159
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 318, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 319, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 320, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 321, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 159u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_2(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 160
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 322, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 325, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 326, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 327, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 160u;
  }
  { // constraint 161
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 323, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 325, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 326, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 327, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 161u;
  }
  { // constraint 162
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 324, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 325, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 326, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 327, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 162u;
  }
  { // constraint 163
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 328, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 331, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 332, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 333, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 163u;
  }
  { // constraint 164
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 329, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 331, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 332, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 333, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 164u;
  }
  { // constraint 165
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 330, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 331, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 332, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 333, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 165u;
  }
  { // constraint 166
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 334, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 337, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 338, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 339, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 166u;
  }
  { // constraint 167
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 335, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 337, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 338, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 339, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 167u;
  }
  { // constraint 168
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 336, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 337, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 338, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 339, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 168u;
  }
  { // constraint 169
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 340, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 343, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 344, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 345, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 169u;
  }
  { // constraint 170
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 341, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 343, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 344, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 345, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 170u;
  }
  { // constraint 171
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 342, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 343, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 344, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 345, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 171u;
  }
  { // constraint 172
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 346, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 349, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 350, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 351, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 172u;
  }
  { // constraint 173
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 347, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 349, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 350, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 351, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 173u;
  }
  { // constraint 174
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 348, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 349, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 350, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 351, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 174u;
  }
  { // constraint 175
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 352, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 355, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 356, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 357, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 175u;
  }
  { // constraint 176
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 353, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 355, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 356, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 357, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 176u;
  }
  { // constraint 177
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 354, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 355, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 356, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 357, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 177u;
  }
  { // constraint 178
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 358, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 361, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 362, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 363, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 178u;
  }
  { // constraint 179
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 359, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 361, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 362, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 363, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 179u;
  }
  { // constraint 180
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 360, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 361, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 362, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 363, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 180u;
  }
  { // constraint 181
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 364, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 367, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 368, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 369, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 181u;
  }
  { // constraint 182
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 365, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 367, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 368, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 369, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 182u;
  }
  { // constraint 183
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 366, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 367, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 368, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 369, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 183u;
  }
  { // constraint 184
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 370, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 373, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 374, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 375, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 184u;
  }
  { // constraint 185
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 371, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 373, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 374, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 375, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 185u;
  }
  { // constraint 186
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 372, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 373, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 374, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 375, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 186u;
  }
  { // constraint 187
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 376, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 379, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 380, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 381, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 187u;
  }
  { // constraint 188
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 377, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 379, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 380, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 381, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 188u;
  }
  { // constraint 189
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 378, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 379, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 380, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 381, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 189u;
  }
  { // constraint 190
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 382, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 385, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 386, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 387, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 190u;
  }
  { // constraint 191
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 383, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 385, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 386, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 387, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 191u;
  }
  { // constraint 192
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 384, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 385, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 386, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 387, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 192u;
  }
  { // constraint 193
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 1, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x99030c2au,0x50f964f9u,0x33cbe473u,0xa81a1a17u,0x591ff444u,0xdaa7af14u,0xe1691a47u,0x2205774eu}}, ldw(wb, 390, bs));
      fr::wide_mac(T, Fr{{0x1993fb98u,0x12890282u,0x0fc92ff1u,0x93220ed9u,0xcf60309du,0xfbe6aed0u,0x39688dfeu,0x287db05bu}}, ldw(wb, 391, bs));
      fr::wide_mac(T, Fr{{0xd9669651u,0x1aed9d1du,0x96aefa48u,0xeac5fadeu,0x34e6347fu,0x3ea16fa4u,0xe183ab12u,0x2c951e29u}}, ldw(wb, 392, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 193u;
  }
  { // constraint 194
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 388, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0xa9b8b22cu,0xdcd469bau,0x6dc8f4e2u,0xd158fac8u,0xb4df49efu,0x5bb77097u,0x3fa6a001u,0x162fbc85u}}, ldw(wb, 390, bs));
      fr::wide_mac(T, Fr{{0xc45863bau,0x53c8b73cu,0x0ad695cbu,0xc88103deu,0xca819ed8u,0x4fbd54e0u,0x97cb8c16u,0x227273b5u}}, ldw(wb, 391, bs));
      fr::wide_mac(T, Fr{{0x2554d8d4u,0x114e3597u,0xc89d2272u,0xa2683a0fu,0xb90bdd20u,0x512fbc5fu,0x2a67c6a8u,0x1899faf8u}}, ldw(wb, 392, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 194u;
  }
  { // constraint 195
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 389, bs);  sc = fr::add(sc, v); }
    { fr::Wide T; fr::wide_zero(T);
      fr::wide_mac(T, Fr{{0x1d7fd8aau,0x77bf9032u,0x1dff7ba3u,0x2b381619u,0xb355d1e4u,0xef60ed09u,0xc58303a0u,0x2acc02beu}}, ldw(wb, 390, bs));
      fr::wide_mac(T, Fr{{0x79401104u,0x2c8bdb42u,0x69b90ee2u,0x0b5e108au,0xa499f3f1u,0x47035603u,0xef1a6040u,0x30208f10u}}, ldw(wb, 391, bs));
      fr::wide_mac(T, Fr{{0xd38d6ed9u,0xf76f11ceu,0x62ac0f24u,0x54d8486fu,0x36ca57b6u,0x6b46e7efu,0x736bcc8eu,0x1de1907bu}}, ldw(wb, 392, bs));
      sc = fr::add(sc, fr::wide_reduce(T, 3)); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 195u;
  }
  { // constraint 196
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 393, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 394, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 196u;
  }
  { // constraint 197
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 394, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 395, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 197u;
  }
  { // constraint 198
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 395, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 393, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 7, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 198u;
  }
  { // constraint 199
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 396, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 397, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 199u;
  }
  { // constraint 200
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 397, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 398, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 200u;
  }
  { // constraint 201
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 398, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 396, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 8, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 201u;
  }
  { // constraint 202
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 399, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 400, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 202u;
  }
  { // constraint 203
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 400, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 401, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 203u;
  }
  { // constraint 204
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 401, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 399, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 9, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 204u;
  }
  { // constraint 205
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 402, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 403, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 205u;
  }
  { // constraint 206
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 403, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 404, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 206u;
  }
  { // constraint 207
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 404, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 402, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 13, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 207u;
  }
  { // constraint 208
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 405, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 406, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 208u;
  }
  { // constraint 209
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 406, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 407, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 209u;
  }
  { // constraint 210
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 407, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 405, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 14, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 210u;
  }
  { // constraint 211
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 408, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 409, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 211u;
  }
  { // constraint 212
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 409, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 410, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 212u;
  }
  { // constraint 213
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 410, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 408, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 15, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 213u;
  }
  { // constraint 214
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 411, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 412, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 214u;
  }
  { // constraint 215
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 412, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 413, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 215u;
  }
  { // constraint 216
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 413, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 411, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 19, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 216u;
  }
  { // constraint 217
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 414, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 415, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 217u;
  }
  { // constraint 218
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 415, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 416, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 218u;
  }
  { // constraint 219
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 416, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 414, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 20, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 219u;
  }
  { // constraint 220
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 417, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 418, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 220u;
  }
  { // constraint 221
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 418, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 419, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 221u;
  }
  { // constraint 222
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 419, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 417, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 21, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 222u;
  }
  { // constraint 223
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 420, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 421, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 223u;
  }
  { // constraint 224
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 421, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 422, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 224u;
  }
  { // constraint 225
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 422, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 420, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 25, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 225u;
  }
  { // constraint 226
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 423, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 424, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 226u;
  }
  { // constraint 227
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 424, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 425, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 227u;
  }
  { // constraint 228
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 425, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 423, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 26, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 228u;
  }
  { // constraint 229
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 426, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 427, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 229u;
  }
  { // constraint 230
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 427, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 428, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 230u;
  }
  { // constraint 231
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 428, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 426, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 27, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 231u;
  }
  { // constraint 232
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 429, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 430, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 232u;
  }
  { // constraint 233
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 430, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 431, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 233u;
  }
  { // constraint 234
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 431, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 429, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 373, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 234u;
  }
  { // constraint 235
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 432, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 433, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 235u;
  }
  { // constraint 236
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 433, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 434, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 236u;
  }
  { // constraint 237
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 434, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 432, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 374, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 237u;
  }
  { // constraint 238
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 435, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 436, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 238u;
  }
  { // constraint 239
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 436, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 437, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 239u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_3(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 240
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 437, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 435, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 375, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 240u;
  }
  { // constraint 241
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 438, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 439, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 241u;
  }
  { // constraint 242
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 439, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 440, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 242u;
  }
  { // constraint 243
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 440, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 438, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 379, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 243u;
  }
  { // constraint 244
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 441, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 442, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 244u;
  }
  { // constraint 245
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 442, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 443, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 245u;
  }
  { // constraint 246
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 443, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 441, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 380, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 246u;
  }
  { // constraint 247
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 444, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 445, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 247u;
  }
  { // constraint 248
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 445, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 446, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 248u;
  }
  { // constraint 249
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 446, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 444, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 381, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 249u;
  }
  { // constraint 250
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 447, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 448, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 250u;
  }
  { // constraint 251
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 448, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 449, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 251u;
  }
  { // constraint 252
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 449, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 447, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 385, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 252u;
  }
  { // constraint 253
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 450, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 451, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 253u;
  }
  { // constraint 254
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 451, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 452, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 254u;
  }
  { // constraint 255
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 452, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 450, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 386, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 255u;
  }
  { // constraint 256
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 453, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 454, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 256u;
  }
  { // constraint 257
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 454, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 455, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 257u;
  }
  { // constraint 258
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 455, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 453, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 387, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 258u;
  }
  { // constraint 259
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 456, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 457, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 259u;
  }
  { // constraint 260
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 457, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 458, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 260u;
  }
  { // constraint 261
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 458, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 456, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 390, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 261u;
  }
  { // constraint 262
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 459, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 460, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 262u;
  }
  { // constraint 263
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 460, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 461, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 263u;
  }
  { // constraint 264
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 461, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 459, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 391, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 264u;
  }
  { // constraint 265
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 462, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 463, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 265u;
  }
  { // constraint 266
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 463, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 464, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 266u;
  }
  { // constraint 267
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 464, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 462, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 392, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 267u;
  }
  { // constraint 268
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 465, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 466, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 268u;
  }
  { // constraint 269
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 466, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 467, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 269u;
  }
  { // constraint 270
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 467, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 465, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 31, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 270u;
  }
  { // constraint 271
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 468, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 469, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 271u;
  }
  { // constraint 272
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 469, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 470, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 272u;
  }
  { // constraint 273
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 470, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 468, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 37, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 273u;
  }
  { // constraint 274
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 471, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 472, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 274u;
  }
  { // constraint 275
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 472, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 473, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 275u;
  }
  { // constraint 276
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 473, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 471, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 43, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 276u;
  }
  { // constraint 277
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 474, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 475, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 277u;
  }
  { // constraint 278
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 475, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 476, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 278u;
  }
  { // constraint 279
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 476, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 474, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 49, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 279u;
  }
  { // constraint 280
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 477, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 478, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 280u;
  }
  { // constraint 281
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 478, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 479, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 281u;
  }
  { // constraint 282
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 479, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 477, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 55, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 282u;
  }
  { // constraint 283
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 480, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 481, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 283u;
  }
  { // constraint 284
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 481, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 482, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 284u;
  }
  { // constraint 285
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 482, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 480, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 61, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 285u;
  }
  { // constraint 286
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 483, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 484, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 286u;
  }
  { // constraint 287
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 484, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 485, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 287u;
  }
  { // constraint 288
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 485, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 483, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 67, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 288u;
  }
  { // constraint 289
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 486, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 487, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 289u;
  }
  { // constraint 290
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 487, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 488, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 290u;
  }
  { // constraint 291
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 488, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 486, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 73, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 291u;
  }
  { // constraint 292
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 489, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 490, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 292u;
  }
  { // constraint 293
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 490, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 491, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 293u;
  }
  { // constraint 294
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 491, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 489, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 79, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 294u;
  }
  { // constraint 295
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 492, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 493, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 295u;
  }
  { // constraint 296
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 493, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 494, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 296u;
  }
  { // constraint 297
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 494, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 492, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 85, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 297u;
  }
  { // constraint 298
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 495, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 496, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 298u;
  }
  { // constraint 299
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 496, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 497, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 299u;
  }
  { // constraint 300
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 497, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 495, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 91, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 300u;
  }
  { // constraint 301
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 498, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 499, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 301u;
  }
  { // constraint 302
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 499, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 500, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 302u;
  }
  { // constraint 303
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 500, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 498, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 97, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 303u;
  }
  { // constraint 304
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 501, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 502, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 304u;
  }
  { // constraint 305
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 502, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 503, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 305u;
  }
  { // constraint 306
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 503, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 501, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 103, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 306u;
  }
  { // constraint 307
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 504, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 505, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 307u;
  }
  { // constraint 308
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 505, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 506, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 308u;
  }
  { // constraint 309
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 506, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 504, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 109, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 309u;
  }
  { // constraint 310
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 507, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 508, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 310u;
  }
  { // constraint 311
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 508, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 509, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 311u;
  }
  { // constraint 312
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 509, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 507, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 115, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 312u;
  }
  { // constraint 313
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 510, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 511, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 313u;
  }
  { // constraint 314
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 511, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 512, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 314u;
  }
  { // constraint 315
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 512, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 510, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 121, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 315u;
  }
  { // constraint 316
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 513, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 514, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 316u;
  }
  { // constraint 317
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 514, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 515, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 317u;
  }
  { // constraint 318
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 515, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 513, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 127, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 318u;
  }
  { // constraint 319
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 516, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 517, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 319u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_4(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 320
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 517, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 518, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 320u;
  }
  { // constraint 321
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 518, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 516, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 133, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 321u;
  }
  { // constraint 322
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 519, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 520, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 322u;
  }
  { // constraint 323
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 520, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 521, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 323u;
  }
  { // constraint 324
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 521, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 519, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 139, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 324u;
  }
  { // constraint 325
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 522, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 523, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 325u;
  }
  { // constraint 326
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 523, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 524, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 326u;
  }
  { // constraint 327
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 524, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 522, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 145, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 327u;
  }
  { // constraint 328
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 525, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 526, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 328u;
  }
  { // constraint 329
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 526, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 527, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 329u;
  }
  { // constraint 330
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 527, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 525, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 151, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 330u;
  }
  { // constraint 331
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 528, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 529, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 331u;
  }
  { // constraint 332
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 529, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 530, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 332u;
  }
  { // constraint 333
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 530, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 528, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 157, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 333u;
  }
  { // constraint 334
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 531, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 532, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 334u;
  }
  { // constraint 335
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 532, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 533, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 335u;
  }
  { // constraint 336
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 533, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 531, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 163, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 336u;
  }
  { // constraint 337
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 534, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 535, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 337u;
  }
  { // constraint 338
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 535, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 536, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 338u;
  }
  { // constraint 339
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 536, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 534, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 169, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 339u;
  }
  { // constraint 340
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 537, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 538, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 340u;
  }
  { // constraint 341
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 538, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 539, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 341u;
  }
  { // constraint 342
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 539, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 537, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 175, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 342u;
  }
  { // constraint 343
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 540, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 541, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 343u;
  }
  { // constraint 344
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 541, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 542, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 344u;
  }
  { // constraint 345
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 542, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 540, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 181, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 345u;
  }
  { // constraint 346
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 543, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 544, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 346u;
  }
  { // constraint 347
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 544, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 545, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 347u;
  }
  { // constraint 348
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 545, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 543, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 187, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 348u;
  }
  { // constraint 349
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 546, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 547, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 349u;
  }
  { // constraint 350
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 547, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 548, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 350u;
  }
  { // constraint 351
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 548, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 546, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 193, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 351u;
  }
  { // constraint 352
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 549, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 550, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 352u;
  }
  { // constraint 353
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 550, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 551, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 353u;
  }
  { // constraint 354
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 551, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 549, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 199, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 354u;
  }
  { // constraint 355
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 552, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 553, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 355u;
  }
  { // constraint 356
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 553, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 554, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 356u;
  }
  { // constraint 357
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 554, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 552, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 205, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 357u;
  }
  { // constraint 358
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 555, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 556, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 358u;
  }
  { // constraint 359
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 556, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 557, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 359u;
  }
  { // constraint 360
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 557, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 555, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 211, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 360u;
  }
  { // constraint 361
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 558, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 559, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 361u;
  }
  { // constraint 362
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 559, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 560, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 362u;
  }
  { // constraint 363
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 560, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 558, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 217, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 363u;
  }
  { // constraint 364
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 561, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 562, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 364u;
  }
  { // constraint 365
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 562, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 563, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 365u;
  }
  { // constraint 366
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 563, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 561, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 223, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 366u;
  }
  { // constraint 367
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 564, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 565, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 367u;
  }
  { // constraint 368
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 565, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 566, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 368u;
  }
  { // constraint 369
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 566, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 564, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 229, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 369u;
  }
  { // constraint 370
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 567, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 568, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 370u;
  }
  { // constraint 371
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 568, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 569, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 371u;
  }
  { // constraint 372
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 569, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 567, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 235, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 372u;
  }
  { // constraint 373
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 570, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 571, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 373u;
  }
  { // constraint 374
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 571, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 572, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 374u;
  }
  { // constraint 375
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 572, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 570, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 241, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 375u;
  }
  { // constraint 376
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 573, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 574, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 376u;
  }
  { // constraint 377
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 574, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 575, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 377u;
  }
  { // constraint 378
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 575, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 573, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 247, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 378u;
  }
  { // constraint 379
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 576, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 577, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 379u;
  }
  { // constraint 380
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 577, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 578, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 380u;
  }
  { // constraint 381
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 578, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 576, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 253, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 381u;
  }
  { // constraint 382
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 579, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 580, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 382u;
  }
  { // constraint 383
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 580, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 581, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 383u;
  }
  { // constraint 384
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 581, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 579, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 259, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 384u;
  }
  { // constraint 385
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 582, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 583, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 385u;
  }
  { // constraint 386
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 583, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 584, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 386u;
  }
  { // constraint 387
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 584, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 582, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 265, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 387u;
  }
  { // constraint 388
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 585, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 586, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 388u;
  }
  { // constraint 389
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 586, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 587, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 389u;
  }
  { // constraint 390
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 587, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 585, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 271, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 390u;
  }
  { // constraint 391
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 588, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 589, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 391u;
  }
  { // constraint 392
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 589, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 590, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 392u;
  }
  { // constraint 393
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 590, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 588, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 277, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 393u;
  }
  { // constraint 394
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 591, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 592, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 394u;
  }
  { // constraint 395
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 592, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 593, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 395u;
  }
  { // constraint 396
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 593, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 591, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 283, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 396u;
  }
  { // constraint 397
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 594, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 595, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 397u;
  }
  { // constraint 398
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 595, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 596, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 398u;
  }
  { // constraint 399
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 596, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 594, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 289, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 399u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_5(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 400
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 597, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 598, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 400u;
  }
  { // constraint 401
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 598, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 599, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 401u;
  }
  { // constraint 402
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 599, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 597, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 295, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 402u;
  }
  { // constraint 403
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 600, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 601, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 403u;
  }
  { // constraint 404
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 601, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 602, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 404u;
  }
  { // constraint 405
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 602, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 600, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 301, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 405u;
  }
  { // constraint 406
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 603, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 604, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 406u;
  }
  { // constraint 407
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 604, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 605, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 407u;
  }
  { // constraint 408
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 605, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 603, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 307, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 408u;
  }
  { // constraint 409
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 606, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 607, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 409u;
  }
  { // constraint 410
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 607, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 608, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 410u;
  }
  { // constraint 411
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 608, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 606, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 313, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 411u;
  }
  { // constraint 412
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 609, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 610, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 412u;
  }
  { // constraint 413
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 610, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 611, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 413u;
  }
  { // constraint 414
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 611, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 609, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 319, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 414u;
  }
  { // constraint 415
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 612, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 613, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 415u;
  }
  { // constraint 416
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 613, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 614, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 416u;
  }
  { // constraint 417
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 614, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 612, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 325, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 417u;
  }
  { // constraint 418
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 615, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 616, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 418u;
  }
  { // constraint 419
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 616, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 617, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 419u;
  }
  { // constraint 420
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 617, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 615, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 331, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 420u;
  }
  { // constraint 421
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 618, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 619, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 421u;
  }
  { // constraint 422
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 619, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 620, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 422u;
  }
  { // constraint 423
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 620, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 618, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 337, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 423u;
  }
  { // constraint 424
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 621, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 622, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 424u;
  }
  { // constraint 425
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 622, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 623, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 425u;
  }
  { // constraint 426
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 623, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 621, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 343, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 426u;
  }
  { // constraint 427
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 624, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 625, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 427u;
  }
  { // constraint 428
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 625, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 626, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 428u;
  }
  { // constraint 429
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 626, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 624, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 349, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 429u;
  }
  { // constraint 430
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 627, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 628, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 430u;
  }
  { // constraint 431
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 628, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 629, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 431u;
  }
  { // constraint 432
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 629, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 627, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 355, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 432u;
  }
  { // constraint 433
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 630, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 631, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 433u;
  }
  { // constraint 434
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 631, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 632, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 434u;
  }
  { // constraint 435
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 632, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 630, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 361, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 435u;
  }
  { // constraint 436
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 633, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 634, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 436u;
  }
  { // constraint 437
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 634, bs);  sa = fr::add(sa, v); }
    Fr prod = fr::mont_sqr(sa);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 635, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 437u;
  }
  { // constraint 438
    Fr sa = fr::zero();
    { Fr v = ldw(wb, 635, bs);  sa = fr::add(sa, v); }
    Fr sb = fr::zero();
    { Fr v = ldw(wb, 633, bs);  sb = fr::add(sb, v); }
    Fr prod = fr::mont_mul(sa, sb);
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 367, bs);  sc = fr::add(sc, v); }
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 438u;
  }
  { // constraint 439
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 393, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8acb57a9u,0xc0263ad0u,0x78946689u,0x1ec286b2u,0x735dc751u,0x90ae2cb1u,0x18e12be7u,0x1a2b8f5au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 439u;
  }
  { // constraint 440
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 2, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 396, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x34da9ce7u,0xe5b644adu,0x9ac3e932u,0xd183a58fu,0xd897a40bu,0xb69407a5u,0x4e4282abu,0x18a46d12u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 440u;
  }
  { // constraint 441
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 3, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 399, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9baf7278u,0x9ee2fbdau,0x2ad4a9f2u,0x78487ccbu,0x85513c23u,0xdf645d0eu,0xd984281cu,0x1823b704u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 441u;
  }
  { // constraint 442
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 4, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 402, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x4fb48d36u,0x2b291040u,0x8d289181u,0xdace988fu,0xe465c204u,0x78c295f2u,0xd2e3359fu,0x15f187cdu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 442u;
  }
  { // constraint 443
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 5, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 405, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xdd38e887u,0xfef065b6u,0x27ad9fa2u,0xe4615dadu,0x7fee6854u,0x7444c2a3u,0xb9b6704bu,0x1c9f0f37u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 443u;
  }
  { // constraint 444
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 6, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 408, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x51cc48abu,0x47d08a27u,0x36ce1d70u,0x94f58078u,0xb7a79a8du,0x0c8f6bc3u,0x35663b79u,0x0cb39153u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 444u;
  }
  { // constraint 445
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 10, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 411, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9cadbc52u,0x6553379bu,0xc5e15d09u,0x09aed4d0u,0x926ed8a4u,0xd8ab94aeu,0x472d7d14u,0x125173efu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 445u;
  }
  { // constraint 446
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 11, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 414, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x12d9b096u,0x618204eau,0xa51c67deu,0x261e033eu,0x3f3f07cau,0xf5813c81u,0x5b669d48u,0x24b95935u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 446u;
  }
  { // constraint 447
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 12, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 417, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x5f307eecu,0xd537af64u,0x57dc427du,0x792c0276u,0x94f8adf9u,0x1e44e00au,0x7ee679fau,0x1095acfdu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 447u;
  }
  { // constraint 448
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 16, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 420, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xa01763e3u,0xaa3784a3u,0xf93709fbu,0x4c22d5aau,0xb54c1d6eu,0x9de84546u,0x22d45693u,0x06c5a688u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 448u;
  }
  { // constraint 449
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 17, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 423, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x2aabeec9u,0x02a402cbu,0xafdb04deu,0x3ced522fu,0xca1175e7u,0x9de9b763u,0x7bb41284u,0x2076dd5cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 449u;
  }
  { // constraint 450
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 18, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 426, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x96ee0922u,0xf785218bu,0xe40ad7aau,0x8197acabu,0xedafc254u,0xc5476d29u,0x11344554u,0x28c7d4a9u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 450u;
  }
  { // constraint 451
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 22, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 465, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xdea4951bu,0x6684a4edu,0x48202c0eu,0xea136e7au,0x503e80b6u,0xac65850du,0x58ae7c75u,0x06e0c46bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 451u;
  }
  { // constraint 452
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 23, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 32, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd40b0609u,0x13b04dd1u,0x147ed777u,0x37a0267cu,0x58d6909au,0xaf2638e5u,0x396cddc3u,0x12414028u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 452u;
  }
  { // constraint 453
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 24, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 33, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc6bf8efcu,0x196f4526u,0x310cf920u,0x001628d7u,0xc55eb410u,0xa0f3f532u,0x643c6497u,0x279b3bdeu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 453u;
  }
  { // constraint 454
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 28, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 468, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6611c3d4u,0xc53074ebu,0xf2e52404u,0x54b0a49eu,0x3a563b4eu,0x8fe469d6u,0xb019d3d6u,0x0b6d9052u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 454u;
  }
  { // constraint 455
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 29, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 38, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9f47149bu,0x61c0d2feu,0xb9f1d9a5u,0x8835954du,0xae1f6b48u,0x9cb15687u,0xe6c16b35u,0x240cfd6du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 455u;
  }
  { // constraint 456
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 30, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 39, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc5d1b9beu,0x4a2f660fu,0xa6ef5d79u,0xdad7e03cu,0x555a39a3u,0xead99735u,0xa59ab6e5u,0x21bfc6e3u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 456u;
  }
  { // constraint 457
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 34, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 471, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1811ff34u,0x28cf54a9u,0x6c2ea35fu,0xaeef865cu,0xba2337b5u,0x7d950791u,0x9205894eu,0x0898c8edu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 457u;
  }
  { // constraint 458
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 35, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 44, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x51a37bb7u,0x8097e8b1u,0xe0c7312fu,0x6d505846u,0xfb9a8798u,0xc2cd651cu,0x3dd4ad4au,0x2b07e14cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 458u;
  }
  { // constraint 459
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 36, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 45, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x7b08b3e0u,0x10008074u,0xffc68477u,0x493ba8f5u,0x4a3cc7a2u,0x7b6874f9u,0x9eb14f10u,0x21778bdfu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 459u;
  }
  { // constraint 460
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 40, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 474, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xca62c1a9u,0x5f8a82cdu,0x8dceba0du,0xdbcba8b4u,0x79dcf42eu,0xbcc43200u,0xa246b2e7u,0x0225bf82u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 460u;
  }
  { // constraint 461
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 41, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 50, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8cf3a14du,0x724a3147u,0x09a10517u,0x22fb909du,0xc3d085f3u,0x9f6200eeu,0x5d914aaeu,0x17ab004du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 461u;
  }
  { // constraint 462
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 42, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 51, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xef53238au,0x716abee3u,0xfcfcb89bu,0x02a89d25u,0x5ae37247u,0x77711732u,0x17caeb29u,0x0fc9a9b7u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 462u;
  }
  { // constraint 463
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 46, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 477, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe8dfdab2u,0xe9320fb7u,0x882c607au,0x19a242c7u,0x902c82a6u,0xa6b03af6u,0xaf969320u,0x085c1d32u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 463u;
  }
  { // constraint 464
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 47, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 56, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf280d437u,0xe1932454u,0x86f6e57du,0x5b95948bu,0xa8fb8b62u,0x3baca9b1u,0xc260c9e6u,0x2d37cfd6u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 464u;
  }
  { // constraint 465
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 48, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 57, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xcf73cda9u,0x84fc7fe7u,0x56406d2du,0x9ebd17afu,0xe6286dadu,0xcd8c80e2u,0x334abc53u,0x080bf8ecu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 465u;
  }
  { // constraint 466
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 52, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 480, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x73678db9u,0x1c839101u,0x738153a8u,0x1810a315u,0xb65ffe67u,0xdf4b340eu,0x878d8b0du,0x06ccc658u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 466u;
  }
  { // constraint 467
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 53, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 62, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x94811ffcu,0xacf77a2fu,0x041eefbbu,0x154f4964u,0x750537c1u,0x785fd4aau,0x8707f6b6u,0x1e5c797cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 467u;
  }
  { // constraint 468
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 54, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 63, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x46bcf6a2u,0x218c84bdu,0x4c64c628u,0x871932c5u,0xa50e59b0u,0x597b00fcu,0x4190c04bu,0x19a9cac0u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 468u;
  }
  { // constraint 469
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 58, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 483, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xa6707cbau,0x70820e0fu,0x9d5d3f85u,0xe3a5bbd3u,0x28a1d689u,0x2c8ce875u,0x30d8fc71u,0x1ec4314bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 469u;
  }
  { // constraint 470
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 59, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 68, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x5b50d27eu,0x7494493eu,0xecf75e4du,0x613d952cu,0x29c803cdu,0x49ea97f5u,0xf124322eu,0x283a3e0au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 470u;
  }
  { // constraint 471
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 60, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 69, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xdfc59232u,0x2ca23288u,0xe0e66eceu,0x5f188463u,0x1c92b66au,0xff585b76u,0x09526a44u,0x2a0fc64bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 471u;
  }
  { // constraint 472
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 64, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 486, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8916e99cu,0x5205187bu,0x5dbca922u,0x78eefa95u,0xc52179abu,0x4ead4d64u,0xe3c95844u,0x19c0a09au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 472u;
  }
  { // constraint 473
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 65, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 74, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8650f54eu,0x1c6253feu,0x586463beu,0xff4f3f04u,0x6c70796au,0x2d37b1dau,0xee4de67au,0x1b390f86u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 473u;
  }
  { // constraint 474
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 66, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 75, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x11b9e8beu,0xc13db6e9u,0x5f47d493u,0xdc60e0f8u,0x8e21a272u,0x236730c5u,0x709dac3du,0x031e67b2u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 474u;
  }
  { // constraint 475
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 70, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 489, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x22136299u,0x30da2db4u,0xe8807bfdu,0xd1951069u,0x1e0cb7a6u,0x850b351au,0x95940088u,0x0eaeace2u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 475u;
  }
  { // constraint 476
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 71, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 80, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe656723au,0x586a9a3au,0x672f4cd1u,0xd38e7513u,0xf98006fau,0x5d3e608du,0x0b28a241u,0x0f39e175u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 476u;
  }
  { // constraint 477
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 72, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 81, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd72eb910u,0xf7ddcf34u,0xb2343342u,0x1762dc83u,0xd0322931u,0x0f6da098u,0x9c213fd4u,0x2104751cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 477u;
  }
  { // constraint 478
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 76, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 492, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x3ae3c824u,0xdb04b1bcu,0xefd7a01du,0x09eda0e7u,0xf2010109u,0x5f6ff8c2u,0x5aa8ef9eu,0x25588f17u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 478u;
  }
  { // constraint 479
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 77, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 86, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xaedd5cf8u,0x28cd0017u,0x4ba64b49u,0x480ad4f1u,0xf6052581u,0x2b142eb1u,0x7fc3e0f2u,0x0e2101a2u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 479u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_6(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 480
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 78, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 87, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x49eab249u,0xa33c2111u,0xbf77414bu,0x566b1d46u,0xa1271b17u,0xffaacc9cu,0xa882095du,0x080a29d0u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 480u;
  }
  { // constraint 481
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 82, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 495, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf5042c9du,0x2937e18du,0x56b93bdau,0x6f5b31c3u,0x51b55008u,0xacd5561du,0x41faf0c3u,0x078bef27u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 481u;
  }
  { // constraint 482
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 83, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 92, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc1021ecbu,0x21e5e106u,0x4fcabfe9u,0x6790e7bbu,0x1a3bb493u,0xc4bef915u,0xaa547ce7u,0x280dfc40u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 482u;
  }
  { // constraint 483
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 84, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 93, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xdd04060eu,0xa297c232u,0xd6ce8720u,0x7634bfacu,0xd22960cfu,0x85ad46e9u,0x210b20bdu,0x302d2786u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 483u;
  }
  { // constraint 484
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 88, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 498, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8dc57dd0u,0xddfdf075u,0x30f569bdu,0xea4d13cbu,0xc3c86d8au,0x0e9471afu,0x471b461du,0x2304415du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 484u;
  }
  { // constraint 485
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 89, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 98, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xadb2a1fdu,0xf37d65a8u,0x8a92f727u,0xafca9a61u,0x7a685f0cu,0x5da7925eu,0xdb5304b9u,0x15600124u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 485u;
  }
  { // constraint 486
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 90, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 99, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x470b1ee8u,0x19688735u,0xa2802857u,0x93c82d03u,0x560166a1u,0xd63a40e4u,0x68c7b9c5u,0x24212770u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 486u;
  }
  { // constraint 487
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 94, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 501, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xda4f6abfu,0x5b83e41du,0x62b008ddu,0x9165901au,0x768deb08u,0xb1f50c32u,0x99b29a79u,0x1eb32f64u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 487u;
  }
  { // constraint 488
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 95, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 104, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xa34dab42u,0x366d04fau,0x7d55b324u,0xd8e59fd6u,0x95a68303u,0x2e788e8eu,0xa822e807u,0x03724befu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 488u;
  }
  { // constraint 489
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 96, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 105, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xcc424fa9u,0xa0d6a32bu,0x80cca383u,0x9fc83091u,0xc557ff48u,0x990cc748u,0xfaae6d04u,0x06b725f8u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 489u;
  }
  { // constraint 490
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 100, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 504, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe01bc5a4u,0x57a6bc52u,0x9cbf5926u,0xe87a0f3au,0xdd0c018cu,0x5885a38cu,0xb2dfe951u,0x28becb5cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 490u;
  }
  { // constraint 491
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 101, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 110, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe9f7c693u,0xe8a21330u,0x4b242c8du,0x09ef5237u,0xde0515b4u,0x32f5a80bu,0x2c374f01u,0x1f6e0a7eu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 491u;
  }
  { // constraint 492
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 102, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 111, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xeb0671ccu,0x1b88e893u,0xb6ece39bu,0x46f9d2a8u,0xc31926e2u,0x2b2d8730u,0x5b18c493u,0x2809eaf7u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 492u;
  }
  { // constraint 493
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 106, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 507, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xae772b2cu,0xee592efeu,0x43aa6d1du,0xff761d5bu,0x680c8918u,0x5083cb75u,0xb469567du,0x08b8d904u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 493u;
  }
  { // constraint 494
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 107, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 116, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe7716654u,0x11413404u,0x9c1cae9eu,0x05ba8bf7u,0x13beee0cu,0x7e8fc05au,0xb2ef15a7u,0x1e2bbe87u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 494u;
  }
  { // constraint 495
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 108, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 117, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd6347a65u,0x96062026u,0xe986752eu,0xf8b6faa3u,0xfa3512beu,0xcc710e9cu,0x24ed8b8bu,0x2cf6d3cau}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 495u;
  }
  { // constraint 496
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 112, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 510, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6a46be39u,0x1737b3bcu,0x646a4426u,0x911599fcu,0x18a07194u,0x3fdb9ce2u,0x8c3bcdf5u,0x28afdb15u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 496u;
  }
  { // constraint 497
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 113, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 122, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x4fbb370fu,0x41ecdb59u,0x7d0e4468u,0xadd07cfcu,0xd817578eu,0x0531923eu,0xd5809b7du,0x2157fa82u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 497u;
  }
  { // constraint 498
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 114, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 123, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb16a991au,0x6c4ecd93u,0xfc88fbd6u,0x3ad4b187u,0x76e30ba4u,0x2e229ef0u,0x51e57867u,0x07a6edabu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 498u;
  }
  { // constraint 499
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 118, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 513, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xcf9ce445u,0x43e3380cu,0x5fe0c4ebu,0x9935be06u,0x028fa22fu,0x4c70f7c9u,0x73c4a99bu,0x106d8492u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 499u;
  }
  { // constraint 500
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 119, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 128, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9847ef94u,0x213b5ebfu,0x1c706a99u,0x9c9ae85bu,0xc592562au,0x5970b28cu,0x9a4753f0u,0x1cd1cc26u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 500u;
  }
  { // constraint 501
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 120, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 129, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf8783761u,0x98d17894u,0x71be3e60u,0x00a56c06u,0x9d13d607u,0x1e965b49u,0xe600001bu,0x22fc1ea1u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 501u;
  }
  { // constraint 502
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 124, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 516, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x3ee37af5u,0xf512fc84u,0xafcaace7u,0xa34f90c6u,0x9e8f6364u,0x298df48du,0xce8a935bu,0x063833eeu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 502u;
  }
  { // constraint 503
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 125, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 134, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbcd1d33au,0xb4ab5b46u,0x45945b00u,0x76079716u,0xc970353du,0x47e58207u,0x565316b0u,0x16f04014u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 503u;
  }
  { // constraint 504
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 126, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 135, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc60d748cu,0x21d174a3u,0x04ee51fdu,0x83fca33fu,0xbae21abau,0x30bdd682u,0xf306cec3u,0x2dc7585bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 504u;
  }
  { // constraint 505
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 130, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 519, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x01ee16e4u,0x419dc944u,0xf9dcb27au,0xca6fd767u,0xcb43c405u,0x9dc9d383u,0x2eb978e1u,0x1647b46bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 505u;
  }
  { // constraint 506
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 131, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 140, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x0b8b0eedu,0x250672b9u,0x065f6691u,0x54776a58u,0x5973afc3u,0xb49cc439u,0x8fb253b0u,0x29626123u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 506u;
  }
  { // constraint 507
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 132, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 141, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x840f7b6au,0x629490f9u,0x7e17fedfu,0x8449a985u,0x1bba219fu,0x28ef3d78u,0xeda5216eu,0x15a2a474u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 507u;
  }
  { // constraint 508
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 136, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 522, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x7b10da2au,0x7498dc11u,0x69830c42u,0xd207e268u,0x5a1332d0u,0x790eae48u,0xc7e2a695u,0x11086942u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 508u;
  }
  { // constraint 509
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 137, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 146, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb4f39aa1u,0xc121a2f6u,0xae7bf804u,0x59aeb060u,0x45a86fa0u,0x2606f619u,0xa15e61d6u,0x1d37cc7eu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 509u;
  }
  { // constraint 510
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 138, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 147, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc632cef0u,0x04a1297du,0xec93f2e9u,0x6e3a81ecu,0x989ff6c4u,0x9ff0e3bdu,0x23fdf508u,0x2e4972cbu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 510u;
  }
  { // constraint 511
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 142, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 525, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb02c9ed1u,0x0fb2cb82u,0x451d1d87u,0x8b471c73u,0x4e3d101bu,0x56d47041u,0xad42be56u,0x1a1b807bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 511u;
  }
  { // constraint 512
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 143, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 152, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6d80308eu,0xa4e62041u,0x0f02dd1bu,0x2d47fd2cu,0xc2f2cfb4u,0x720ed2ceu,0xd42d741au,0x281cbb54u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 512u;
  }
  { // constraint 513
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 144, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 153, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x65d72395u,0x31e0622fu,0x92826b6bu,0x8fa98545u,0xea3a0d50u,0x10c16aa8u,0x3f7c582fu,0x16ceb831u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 513u;
  }
  { // constraint 514
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 148, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 528, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xcbd83d2fu,0xb573fa82u,0x94308d0cu,0x0771902eu,0x50f7dba7u,0x9ac80e39u,0x6d782d8fu,0x0613c37fu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 514u;
  }
  { // constraint 515
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 149, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 158, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x123ae981u,0x26cd2d20u,0x71b8da3bu,0x1a263fdfu,0x6e4081b9u,0xf9017bfcu,0x8726a48cu,0x2fcb907cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 515u;
  }
  { // constraint 516
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 150, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 159, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xfcdbc1a1u,0xf30abf7eu,0x94ad1d12u,0xe31349deu,0xba279a49u,0xf90b6c16u,0xda0cfaa5u,0x08c51993u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 516u;
  }
  { // constraint 517
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 154, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 531, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb3c5597fu,0x2244f5adu,0x375ea149u,0x81568e81u,0xd4b46bcdu,0x74038eafu,0xe1d60255u,0x01d83693u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 517u;
  }
  { // constraint 518
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 155, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 164, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb940a2f6u,0xaa4a2c7fu,0x1242f7d2u,0x1209ed00u,0xa5e2a5c3u,0x6e8b6c3au,0x0690ca0bu,0x102e7169u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 518u;
  }
  { // constraint 519
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 156, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 165, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x61aeab96u,0x33cb6717u,0x0f074385u,0xd7b526f7u,0xdcc24e9cu,0x9e36f5ebu,0x78e7ab2cu,0x2400d0aau}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 519u;
  }
  { // constraint 520
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 160, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 534, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd1590ad0u,0xebe3f12au,0x447c6d15u,0x06a3801au,0x987c84a4u,0x456bc6bcu,0xe5f52a3cu,0x0f450714u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 520u;
  }
  { // constraint 521
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 161, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 170, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9fc9e831u,0xc55e6a0du,0xeed7db32u,0x417916b3u,0xabce53d9u,0x47d1ecbbu,0xb5bb54f8u,0x265adbdau}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 521u;
  }
  { // constraint 522
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 162, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 171, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf3cbe7a5u,0x40266b8au,0x2c386494u,0xb19c3b1bu,0xb96a635bu,0xbb3e3246u,0x590a04d2u,0x1e574706u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 522u;
  }
  { // constraint 523
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 166, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 537, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x04dd7ff5u,0x55bde535u,0xf7913332u,0x70680098u,0xf4aa3ca7u,0x56be63d2u,0x4f41c680u,0x15795b8eu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 523u;
  }
  { // constraint 524
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 167, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 176, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6336dc87u,0xa393ee92u,0x48678fdau,0xd4d2750du,0xf3a1c099u,0x776a26d2u,0xe87e9055u,0x0de9555bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 524u;
  }
  { // constraint 525
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 168, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 177, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd24378d8u,0x09583381u,0xa0de560au,0x6ba42a94u,0xf3fd4cd6u,0xcfe1fd06u,0xd6626e60u,0x037ca9a7u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 525u;
  }
  { // constraint 526
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 172, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 540, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x13c50b9du,0xf3ad5503u,0x184967e6u,0xb95f78c6u,0xa9676079u,0x82c1d3b6u,0x001abeeeu,0x099c4668u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 526u;
  }
  { // constraint 527
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 173, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 182, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x374ad1cau,0xcb58a2c4u,0x91bdf0d4u,0x386cbdecu,0xed2813e2u,0xf0d49966u,0x56c47a0au,0x30540dc3u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 527u;
  }
  { // constraint 528
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 174, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 183, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1d008bdbu,0x2cfb1f93u,0xccbcdc90u,0x5dd49031u,0xe77357eau,0xd4c456b7u,0xc6b5dcf7u,0x0ed0c359u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 528u;
  }
  { // constraint 529
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 178, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 543, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x4de06d6au,0x6e2da872u,0xcf3c694du,0xa21d6ed4u,0x30ac730du,0x3e92d671u,0x34c85f5fu,0x2c947332u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 529u;
  }
  { // constraint 530
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 179, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 188, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1962aa05u,0xa74f59f2u,0x64a6fee2u,0xa05d9a4cu,0xe1d4c326u,0x8b28629eu,0x502fb7b0u,0x15d39993u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 530u;
  }
  { // constraint 531
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 180, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 189, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x2f141815u,0xcf488f44u,0xf5265d8du,0x1ec278b9u,0xed1e0d18u,0x6b8f68eau,0x5c17bcedu,0x16315873u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 531u;
  }
  { // constraint 532
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 184, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 546, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb4643558u,0xbcd97ddcu,0x0481d241u,0x0fb26139u,0x2fd971ebu,0x19113ed2u,0x2d13d500u,0x1f125037u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 532u;
  }
  { // constraint 533
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 185, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 194, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd353fbd1u,0x386543ebu,0x3599e798u,0x0114a02cu,0xe936f5c9u,0xcaaedc6cu,0xdcf5e70cu,0x29278574u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 533u;
  }
  { // constraint 534
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 186, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 195, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbb73401du,0xe695a995u,0x324cdc07u,0x585d9384u,0x90cf688fu,0xdb075079u,0x77cf9017u,0x299ad888u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 534u;
  }
  { // constraint 535
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 190, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 549, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd7b36fddu,0x09e6891fu,0xd8757270u,0xa74c4d0eu,0xe4388c60u,0x072b0865u,0x54df8ddcu,0x161b4db4u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 535u;
  }
  { // constraint 536
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 191, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 200, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xa2ae4a41u,0x2ad79bc2u,0xa0fbb957u,0x83ef062du,0x837b40c4u,0x062aadaeu,0x216270bfu,0x094a1fd7u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 536u;
  }
  { // constraint 537
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 192, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 201, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x59be8eceu,0x8332e1b7u,0xdd312e83u,0x0b821fb0u,0x0dc7601cu,0x8310b760u,0xc8638ea6u,0x2591c72au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 537u;
  }
  { // constraint 538
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 196, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 552, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf2f082d8u,0xe533e4bdu,0x8f111cacu,0xaa931c1cu,0xaedda17au,0x000b956bu,0xa071673bu,0x254bff67u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 538u;
  }
  { // constraint 539
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 197, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 206, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x3e2aba70u,0x5df85fdfu,0xc88b1ff7u,0x0edf591fu,0x8d723bebu,0xd775ec3fu,0xb4305bd2u,0x140f18d5u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 539u;
  }
  { // constraint 540
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 198, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 207, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x90907829u,0xdb1de44du,0xb038a739u,0x84925fc5u,0x31d5893cu,0x5bd6f42au,0x49754376u,0x09768bdeu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 540u;
  }
  { // constraint 541
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 202, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 555, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x5b7498cfu,0x8ed291d0u,0x5ed794dcu,0x8c1ddc71u,0x26c161f2u,0x93806c7eu,0xf4440183u,0x2e5e13f5u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 541u;
  }
  { // constraint 542
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 203, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 212, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1dddba51u,0x68d5395eu,0x24964223u,0x29adb6aeu,0x8dd38324u,0x574a49b8u,0xa8585dd4u,0x04368bf6u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 542u;
  }
  { // constraint 543
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 204, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 213, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6573766cu,0x525897fbu,0x6d9e160eu,0x4954ec9bu,0x75ad4184u,0x59108cfau,0xbf5a1267u,0x223a01f5u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 543u;
  }
  { // constraint 544
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 208, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 558, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x37832ae9u,0xd0ec5653u,0x4c2e31adu,0xc698c768u,0x89a825bdu,0xe1648d1fu,0x7fefc2eeu,0x0f482e93u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 544u;
  }
  { // constraint 545
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 209, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 218, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xa0ab366fu,0x6cfe82cdu,0x8233c4b0u,0x8eb16d1cu,0x31c2aa48u,0x01dbb2ddu,0x4fd28372u,0x2ddc3728u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 545u;
  }
  { // constraint 546
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 210, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 219, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd5b52e3cu,0x36f36c33u,0xa44887ddu,0xc155009bu,0x2d0f5c30u,0x94dd1404u,0xdab17f59u,0x20eaaa96u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 546u;
  }
  { // constraint 547
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 214, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 561, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x60ce43efu,0xba520788u,0x4c17da74u,0xcbaa939eu,0xbda745adu,0x411ddf3du,0x66b2678au,0x1c213b51u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 547u;
  }
  { // constraint 548
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 215, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 224, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1c4b79b7u,0x3deeb2c9u,0x7362d081u,0x456c14abu,0xa71eb61du,0x569786c0u,0x65cb093au,0x170c47c5u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 548u;
  }
  { // constraint 549
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 216, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 225, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe34210ffu,0xcd1af99eu,0x1f4e2f9cu,0x32fdfdb1u,0x4198ec8eu,0xabf0637eu,0x95ab1770u,0x1ee98fbfu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 549u;
  }
  { // constraint 550
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 220, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 564, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbd1ca6ddu,0xd1b22802u,0x4a156a2bu,0x9a9acc39u,0xde4594d7u,0x0d4a4250u,0x82113f7fu,0x250307e5u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 550u;
  }
  { // constraint 551
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 221, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 230, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x60e41dd7u,0x50640e19u,0x0441f872u,0xe5c28e9fu,0x5b200f83u,0xa366ada3u,0xd2d5bf65u,0x1cd5c163u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 551u;
  }
  { // constraint 552
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 222, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 231, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x2644903au,0x628dbe60u,0x43d75404u,0x4f6a7803u,0xfea314d0u,0xe41c57f3u,0xe91668c5u,0x1f1de391u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 552u;
  }
  { // constraint 553
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 226, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 567, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb6adb78du,0x172468b8u,0x2ebc9083u,0x2eae8ce4u,0xd58581f0u,0x05bf7a53u,0x30295b95u,0x0ab779fau}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 553u;
  }
  { // constraint 554
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 227, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 236, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x3be23119u,0xb5258f21u,0xf527d9efu,0xc7c4e102u,0xc3f88032u,0xd2f56525u,0x9fca7b8bu,0x2fcfde1du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 554u;
  }
  { // constraint 555
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 228, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 237, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x2859be76u,0x3dc160d7u,0xf435aa5au,0x00309eecu,0x0d49984du,0x6464fcfdu,0xf81e1ec4u,0x23ea3c4fu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 555u;
  }
  { // constraint 556
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 232, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 570, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb9fe75b3u,0x1bf73df7u,0xa8f56a04u,0x6b1b8386u,0x5bdf2079u,0xbf4e347cu,0xf6130409u,0x1dd6c6c6u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 556u;
  }
  { // constraint 557
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 233, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 242, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf89ad2edu,0x6862317bu,0xcef8d229u,0xcf76990du,0x34207000u,0x45c0ed50u,0x27c1c289u,0x20dea83eu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 557u;
  }
  { // constraint 558
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 234, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 243, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x0cfd1525u,0x60b21593u,0x3d3b800du,0xab10246bu,0xcd3db84au,0x5c7bf4aeu,0xb85ff8edu,0x0597abfeu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 558u;
  }
  { // constraint 559
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 238, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 573, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xdf6bbd9du,0x846f77a6u,0x91da514eu,0xd974c3a8u,0x2d3926e0u,0x1d6863ccu,0xa8c29958u,0x26dfd5a1u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 559u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" __global__ void __launch_bounds__(128, 4) spec_check_7(const uint4 *store, uint64_t bs, uint64_t B, uint32_t *first_bad) {
  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;
  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;
  { // constraint 560
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 239, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 248, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x39d39271u,0x0271c5e5u,0x1c8782ebu,0xca4987f2u,0x8e1e6c17u,0x1f3bf91du,0x983a6125u,0x29fd75ecu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 560u;
  }
  { // constraint 561
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 240, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 249, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd6629665u,0xc02e2b15u,0xb57c110fu,0xf5482497u,0x6763f8f1u,0xa3aa26d9u,0x402166b1u,0x03ae8574u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 561u;
  }
  { // constraint 562
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 244, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 576, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9b59743du,0xe6dc11e5u,0x4558c99eu,0xd8894b02u,0x7a321dadu,0x39116fe9u,0x064b2246u,0x033c3f52u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 562u;
  }
  { // constraint 563
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 245, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 254, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x82525dc1u,0x70f210c9u,0xe7d28c40u,0xa270a5fbu,0x9967c72du,0x50716c8du,0xcb8487abu,0x061d3d71u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 563u;
  }
  { // constraint 564
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 246, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 255, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbda6eb47u,0x4679590bu,0xb4c2f7e8u,0x06b71ca7u,0xb6d52532u,0x40f40b15u,0x337dc41eu,0x2e2c5f73u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 564u;
  }
  { // constraint 565
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 250, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 579, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe2b754a7u,0xaec53b70u,0x0a3ba310u,0x7bd2e953u,0x1b0747e2u,0xca42342au,0xce85b43du,0x07faea94u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 565u;
  }
  { // constraint 566
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 251, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 260, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb8abaeffu,0x6e22ae47u,0x5511b499u,0x9a43c741u,0xb1146a00u,0x8fe665a6u,0xf7ca05a9u,0x1fbea4c8u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 566u;
  }
  { // constraint 567
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 252, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 261, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x483e03d0u,0x8c8b204bu,0x8d25f61eu,0x968ae8c0u,0x4e02bf5cu,0xf7fdc287u,0x49df7968u,0x2b3aca28u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 567u;
  }
  { // constraint 568
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 256, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 582, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x058dc15au,0x6d24937au,0x321232d4u,0x76d4f0eeu,0xc4a5b439u,0xea4320c2u,0xd5bbc593u,0x2f7184b2u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 568u;
  }
  { // constraint 569
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 257, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 266, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb6cba9e6u,0xbba83cafu,0xe0435be1u,0x395e23ceu,0x9e028ee7u,0x2a3c9669u,0xc723f858u,0x14529d3au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 569u;
  }
  { // constraint 570
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 258, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 267, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd4ad58f4u,0x5058f271u,0x23a27fdcu,0xd09fcc9bu,0xbf33ef02u,0x3e6f5fabu,0x936c8e27u,0x084ae404u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 570u;
  }
  { // constraint 571
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 262, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 585, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb22db14bu,0x3d933787u,0x7b0920f4u,0xf030eac3u,0xf101b019u,0x319d7305u,0x5a6b887fu,0x040cfe7cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 571u;
  }
  { // constraint 572
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 263, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 272, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1d94c3cau,0x009135cau,0x21941aeau,0x067368ccu,0x479b3ac5u,0xf5e13b9cu,0xc0ba67a3u,0x0c840317u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 572u;
  }
  { // constraint 573
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 264, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 273, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd3b69f10u,0xa5140513u,0xcf5d46d3u,0x97ee89a7u,0xd9f59ab6u,0x3e00d48du,0xe12b1615u,0x27bd1f50u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 573u;
  }
  { // constraint 574
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 268, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 588, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x9aa9edc8u,0xb569b363u,0xc30fbfcdu,0x61ad904fu,0x42e6c003u,0x7c3115b4u,0x4bb567beu,0x2e59fb6bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 574u;
  }
  { // constraint 575
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 269, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 278, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x10bec791u,0x812f090au,0x0f7cb0ebu,0xbeafd0e4u,0x789cf61bu,0x3297b885u,0x13f588f8u,0x1f107f0eu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 575u;
  }
  { // constraint 576
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 270, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 279, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd4e05304u,0x701ab29fu,0xab636fb8u,0xdf67d650u,0x5bd45192u,0xdf2ba3b2u,0x2ae76938u,0x04d175dcu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 576u;
  }
  { // constraint 577
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 274, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 591, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd9dcd181u,0xfbfe9175u,0x85a7e3aau,0x727bab0fu,0xf2e16f05u,0x1bdac834u,0x5c969e3du,0x21604e51u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 577u;
  }
  { // constraint 578
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 275, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 284, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb68f69ffu,0x84a51e25u,0x82a20f31u,0x455aa9dbu,0xd736b3d9u,0x69f0bfaeu,0x7c8351f8u,0x10976350u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 578u;
  }
  { // constraint 579
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 276, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 285, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xac85c728u,0x790a4334u,0x35fc0f16u,0x31c61b01u,0xb6e8f91bu,0x42edcca5u,0x8d8f2802u,0x1436b850u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 579u;
  }
  { // constraint 580
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 280, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 594, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x56b3319cu,0x0517ee02u,0xfb34f9b8u,0x22a1b396u,0xeed395fbu,0x344c54e8u,0x94922c70u,0x0f4d894eu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 580u;
  }
  { // constraint 581
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 281, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 290, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x031f9a1eu,0xc52538c3u,0xd9a5d097u,0x66937f42u,0x968489e8u,0xc76fa3a1u,0x957dae7bu,0x20422cd0u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 581u;
  }
  { // constraint 582
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 282, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 291, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf2f86c3bu,0x7ca8aec0u,0xd24823e7u,0x51aefedfu,0x2ae09d7fu,0x7339bc90u,0x7774a841u,0x04b2ac9bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 582u;
  }
  { // constraint 583
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 286, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 597, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x74ad1c38u,0xd89ce106u,0x9a2d695cu,0x2527ecd9u,0xf2a41f99u,0x4018fa71u,0x9a8be9a0u,0x2aaa2153u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 583u;
  }
  { // constraint 584
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 287, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 296, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc435cb2bu,0xad39a58du,0x41676f42u,0x0d3c7bb6u,0x83db8836u,0x0abc7759u,0xc67dcce1u,0x140b21c3u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 584u;
  }
  { // constraint 585
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 288, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 297, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xff44b6cfu,0x9426de4fu,0x6ddee484u,0xbdbd5167u,0x20e562ceu,0xb0afcd85u,0x490a8f14u,0x032e666du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 585u;
  }
  { // constraint 586
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 292, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 600, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf2503545u,0x67bd3253u,0x46a11dbcu,0x33569e10u,0xb802cad3u,0x597e8e1bu,0x385ea5a6u,0x28678696u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 586u;
  }
  { // constraint 587
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 293, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 302, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xa2e47c11u,0x5fd1b493u,0x4521a026u,0x1ea0ac6du,0x75c3841cu,0x14addda6u,0x143b6735u,0x02b7797cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 587u;
  }
  { // constraint 588
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 294, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 303, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x0d7e1b18u,0x01f9dcc7u,0x2d1a6555u,0x4d68e26eu,0xea6f1cbeu,0x1bb75daeu,0x6f620796u,0x2aa1d621u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 588u;
  }
  { // constraint 589
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 298, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 603, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x89e22c12u,0xb82aee80u,0x97d109acu,0x85124d06u,0xeca5a891u,0x700e9707u,0x85fd63fbu,0x115fa36bu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 589u;
  }
  { // constraint 590
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 299, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 308, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x15e12b29u,0x3a5a79fau,0xbe6345dau,0xea18d489u,0xc507f155u,0xbb46007bu,0x514c5d72u,0x155b8381u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 590u;
  }
  { // constraint 591
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 300, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 309, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xba12b226u,0xbc6f7e27u,0x0beee170u,0x564f6484u,0x949d8ce5u,0x7ae15005u,0x34f4c14fu,0x2c2f4112u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 591u;
  }
  { // constraint 592
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 304, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 606, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xc2027d74u,0x6e4a6bbeu,0xb1f30527u,0x575ef280u,0xcc50c128u,0x76cb8624u,0x601e2cc4u,0x225a6f95u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 592u;
  }
  { // constraint 593
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 305, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 314, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x88cf3a33u,0x079b4fe6u,0x9f67208au,0xcbb814ccu,0x7710e44bu,0xe2d3279fu,0xc018244bu,0x1c82c4bdu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 593u;
  }
  { // constraint 594
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 306, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 315, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xfcfb5d92u,0x9dadb6b1u,0xa6fefe68u,0x5e0020b4u,0x80ac8414u,0x92757491u,0xbbac73e3u,0x2694cbfcu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 594u;
  }
  { // constraint 595
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 310, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 609, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x29c5ab28u,0xca676f91u,0x3f188869u,0x0d712c05u,0x2975e0d4u,0x3c108737u,0x28aec795u,0x2d4d083du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 595u;
  }
  { // constraint 596
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 311, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 320, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbe76f729u,0xf4cb1adeu,0xde0631f5u,0xc6acde99u,0x61b48aa3u,0x3e597719u,0x7af34ad0u,0x06218c1au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 596u;
  }
  { // constraint 597
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 312, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 321, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6247a414u,0xb08f2f24u,0xe89805aeu,0x30c97b4bu,0xd61379a2u,0xcba289c7u,0xdf96d439u,0x2b62c865u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 597u;
  }
  { // constraint 598
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 316, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 612, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x549cc417u,0x1025fe3eu,0x42f0ae3fu,0xce77af1cu,0x00cfa9a0u,0xa4afef78u,0xdacde8c8u,0x1a69b70au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 598u;
  }
  { // constraint 599
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 317, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 326, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x731a999fu,0x70967ec7u,0x26c8a8f9u,0xe1fc9228u,0xeeaa67beu,0xa6347c23u,0x8c03c3fcu,0x25fafea4u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 599u;
  }
  { // constraint 600
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 318, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 327, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6169e24cu,0x17503cb4u,0x886e11d9u,0x47af7010u,0x2175bfd2u,0xd8a685f6u,0x3e543e3cu,0x21db8e13u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 600u;
  }
  { // constraint 601
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 322, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 615, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1ea6b12fu,0x3db6f46du,0xe658a79eu,0x317af8b3u,0x1961bd43u,0x2fe7e983u,0x04d05837u,0x2ae7aba4u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 601u;
  }
  { // constraint 602
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 323, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 332, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x72e29679u,0x87953d70u,0xb99347a3u,0xe78b593du,0xaafa7f33u,0x7f77ad1du,0x30e043f8u,0x0d308161u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 602u;
  }
  { // constraint 603
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 324, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 333, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x88cdf74cu,0x919ff274u,0x86af8438u,0x13e4f468u,0x9ca4febeu,0x437adf3du,0x60e81b3au,0x03babb71u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 603u;
  }
  { // constraint 604
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 328, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 618, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbc5b4130u,0xf2b019f1u,0xa17964e4u,0x9da5cd11u,0x1bf13b8au,0x966aa72eu,0xed36ef84u,0x2207bfabu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 604u;
  }
  { // constraint 605
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 329, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 338, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x67123fbau,0x19963f96u,0xb32c6467u,0xccacf45cu,0x2af8a80du,0x63c8168cu,0x90b4e1f6u,0x2abafffdu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 605u;
  }
  { // constraint 606
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 330, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 339, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf713246fu,0x18c54ad0u,0x9ad8b084u,0x086a1b8fu,0x4b0ab661u,0xa524acbdu,0x99a11757u,0x09fc1ae2u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 606u;
  }
  { // constraint 607
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 334, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 621, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8aeda764u,0x0eab8821u,0xe8869e91u,0x0df57caeu,0x19feae1bu,0xda47c96du,0x93ddc328u,0x16a73ea7u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 607u;
  }
  { // constraint 608
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 335, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 344, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x0393f89fu,0x3f653e43u,0x1c0dbc16u,0x035d43dau,0x65e00bb2u,0xef37204au,0x54e63d64u,0x0930b7cbu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 608u;
  }
  { // constraint 609
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 336, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 345, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x98978273u,0x53286968u,0x59563a32u,0xa81f6bb7u,0xefc74255u,0x5e9be3cbu,0xd4fea0d9u,0x07a0525du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 609u;
  }
  { // constraint 610
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 340, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 624, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb2bcbdcfu,0x4df52677u,0x9c7c5b5au,0xe100b76eu,0xbd13d1aau,0x0a6ac1efu,0xc26c11b9u,0x040ac1dcu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 610u;
  }
  { // constraint 611
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 341, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 350, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xe210986au,0x4029f57du,0xd0765913u,0x63649d8eu,0xc028c920u,0xea688f8bu,0x8f6223f4u,0x297582a2u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 611u;
  }
  { // constraint 612
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 342, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 351, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb3942e04u,0x4078433bu,0x02e53f1au,0x1a7626a8u,0xd3b27b20u,0xed42743cu,0x9a26e2ffu,0x24a64b02u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 612u;
  }
  { // constraint 613
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 346, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 627, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x7f5aff3au,0xc6f2ae9eu,0xfed90972u,0x526b4b1cu,0x9f25faffu,0x85abd2e3u,0x7d00d319u,0x184af728u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 613u;
  }
  { // constraint 614
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 347, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 356, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xfdeec23fu,0xd4640dbbu,0x9754dbc4u,0x8c8671d1u,0x3204a316u,0x5e98ca3bu,0xd7c581f2u,0x0a3f6295u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 614u;
  }
  { // constraint 615
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 348, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 357, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf35994b8u,0x323c27b5u,0x2bcd84aeu,0xb1a81c1fu,0x01c17e44u,0x332fc4eeu,0xc77b6c3bu,0x13e6bcccu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 615u;
  }
  { // constraint 616
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 352, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 630, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x6eba1054u,0x98fc1236u,0x98fd8a9du,0x5a16977du,0xcf6e09bdu,0x81b938c6u,0x9d6d332bu,0x1de3785du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 616u;
  }
  { // constraint 617
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 353, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 362, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x84ab973bu,0x893302c0u,0xf41aeaf0u,0xdea4685cu,0x926e39b0u,0xbc91f666u,0x215de804u,0x25322ea0u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 617u;
  }
  { // constraint 618
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 354, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 363, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xf4e48567u,0xd012c7bbu,0x94f5098cu,0x2da2d4dau,0xb0dc6954u,0x43d91c78u,0xc32558c9u,0x15361f30u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 618u;
  }
  { // constraint 619
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 358, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 633, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x1487740fu,0x8f329d8eu,0xe8eade39u,0xdc2d8306u,0xfb565449u,0x7c4702f7u,0xf47fa04du,0x172ec3cau}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 619u;
  }
  { // constraint 620
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 359, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 368, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x3191bd2cu,0x2982d30au,0xb5abe617u,0x6fde59beu,0xad6a2f52u,0x4a9c41d5u,0x5cfc0cc0u,0x226b8d45u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 620u;
  }
  { // constraint 621
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 360, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 369, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x605af186u,0x17cbd25cu,0xf342d07eu,0xf4ecde74u,0x96673474u,0x7606ded6u,0x96432a61u,0x1c6d25bau}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 621u;
  }
  { // constraint 622
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 364, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 429, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x325da0cdu,0xeb12ed05u,0xb7af6a94u,0xadd67e10u,0xf28f10cdu,0xf469c7a7u,0xce2d8ed4u,0x0795389au}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 622u;
  }
  { // constraint 623
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 365, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 432, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x8f1d97adu,0x7646be7du,0xbbd1adebu,0x809b0841u,0xf2747a0eu,0x12c679f6u,0xfacc5063u,0x239c1a10u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 623u;
  }
  { // constraint 624
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 366, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 435, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd30f7487u,0xe93c160bu,0xf3086596u,0xa2eeb191u,0xb5d1bd7cu,0x6354693bu,0xb61e87a6u,0x0dfc68f9u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 624u;
  }
  { // constraint 625
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 370, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 438, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x7c2bfceeu,0xdbe12c8du,0x73661ad5u,0xf127258au,0xe8c8c398u,0x6f9ce406u,0xee5bc5dcu,0x1af014f3u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 625u;
  }
  { // constraint 626
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 371, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 441, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xd6c31975u,0xabdd91acu,0x5f6eb14bu,0x79ac0babu,0x87e47cfeu,0xb4c5d2bfu,0xb6a2899du,0x19c6a97du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 626u;
  }
  { // constraint 627
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 372, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 444, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbb97c448u,0x7baa1f4eu,0x222fe57eu,0x866e3b01u,0xdda7ae1du,0xf364e01eu,0xf8013d8eu,0x200f2f0du}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 627u;
  }
  { // constraint 628
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 376, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 447, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xb05c2027u,0x816b03a1u,0x6a6b3781u,0x2075d0abu,0x8ad9c530u,0xc923824du,0xbbc2f960u,0x19cbc575u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 628u;
  }
  { // constraint 629
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 377, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 450, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x5024b6e1u,0x3aadfbadu,0xae6445cfu,0xb2bc4f75u,0xd188b174u,0x73eed084u,0x32ee023bu,0x13db720cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 629u;
  }
  { // constraint 630
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 378, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 453, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xce62fee7u,0xbf0afaf2u,0x3a619c7du,0xbbd1c881u,0xc4088c68u,0xb1ea2e3fu,0x34da589cu,0x2323f7d4u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 630u;
  }
  { // constraint 631
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 382, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 456, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xbc8b9b71u,0xde3a90c8u,0x0c7c5d55u,0x4fdc86bdu,0xc4c92ae0u,0x5ed2afd4u,0x860c941cu,0x16ca9c48u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 631u;
  }
  { // constraint 632
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 383, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 459, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0x02f0aa16u,0x3a50b62fu,0xc7e4d4fcu,0xb63f2fb3u,0xe9ff53f1u,0xd201182fu,0xa5fb757eu,0x059f3205u}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 632u;
  }
  { // constraint 633
    Fr prod = fr::zero();
    Fr sc = fr::zero();
    { Fr v = ldw(wb, 384, bs);  sc = fr::sub(sc, v); }
    { Fr v = ldw(wb, 462, bs);  sc = fr::add(sc, v); }
    sc = fr::add(sc, Fr{{0xfc79baddu,0xf5c2d5dcu,0x5284cdfbu,0x0ff94163u,0x4036e130u,0x3df7e7efu,0x2e483729u,0x2235cf2cu}});
    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = 633u;
  }
  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);
}
extern "C" int spec_check_launch(const void *store, uint64_t bs, uint64_t B, void *first_bad, void *stream) {
  cudaMemsetAsync(first_bad, 0xff, B * 4, (cudaStream_t)stream);
  spec_check_0<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_1<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_2<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_3<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_4<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_5<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_6<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  spec_check_7<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, (uint32_t *)first_bad);
  return (int)cudaGetLastError(); }
