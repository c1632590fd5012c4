// roots of unity exp(-2*pi*i*k/32), 25 significant digits; tools/check_roots32.py re-derives and checks every entry
#pragma once
namespace offtb {
template <int K32> struct Root32;
template <> struct Root32<0> { static constexpr double re = 1.0; static constexpr double im = 0.0; };
template <> struct Root32<1> { static constexpr double re = 0.9807852804032304491261822; static constexpr double im = -0.1950903220161282678482849; };
template <> struct Root32<2> { static constexpr double re = 0.9238795325112867561281832; static constexpr double im = -0.3826834323650897717284600; };
template <> struct Root32<3> { static constexpr double re = 0.8314696123025452370787884; static constexpr double im = -0.5555702330196022247428308; };
template <> struct Root32<4> { static constexpr double re = 0.7071067811865475244008444; static constexpr double im = -0.7071067811865475244008444; };
template <> struct Root32<5> { static constexpr double re = 0.5555702330196022247428308; static constexpr double im = -0.8314696123025452370787884; };
template <> struct Root32<6> { static constexpr double re = 0.3826834323650897717284600; static constexpr double im = -0.9238795325112867561281832; };
template <> struct Root32<7> { static constexpr double re = 0.1950903220161282678482849; static constexpr double im = -0.9807852804032304491261822; };
template <> struct Root32<8> { static constexpr double re = 0.0; static constexpr double im = -1.0; };
template <> struct Root32<9> { static constexpr double re = -0.1950903220161282678482849; static constexpr double im = -0.9807852804032304491261822; };
template <> struct Root32<10> { static constexpr double re = -0.3826834323650897717284600; static constexpr double im = -0.9238795325112867561281832; };
template <> struct Root32<11> { static constexpr double re = -0.5555702330196022247428308; static constexpr double im = -0.8314696123025452370787884; };
template <> struct Root32<12> { static constexpr double re = -0.7071067811865475244008444; static constexpr double im = -0.7071067811865475244008444; };
template <> struct Root32<13> { static constexpr double re = -0.8314696123025452370787884; static constexpr double im = -0.5555702330196022247428308; };
template <> struct Root32<14> { static constexpr double re = -0.9238795325112867561281832; static constexpr double im = -0.3826834323650897717284600; };
template <> struct Root32<15> { static constexpr double re = -0.9807852804032304491261822; static constexpr double im = -0.1950903220161282678482849; };
template <> struct Root32<16> { static constexpr double re = -1.0; static constexpr double im = 0.0; };
template <> struct Root32<17> { static constexpr double re = -0.9807852804032304491261822; static constexpr double im = 0.1950903220161282678482849; };
template <> struct Root32<18> { static constexpr double re = -0.9238795325112867561281832; static constexpr double im = 0.3826834323650897717284600; };
template <> struct Root32<19> { static constexpr double re = -0.8314696123025452370787884; static constexpr double im = 0.5555702330196022247428308; };
template <> struct Root32<20> { static constexpr double re = -0.7071067811865475244008444; static constexpr double im = 0.7071067811865475244008444; };
template <> struct Root32<21> { static constexpr double re = -0.5555702330196022247428308; static constexpr double im = 0.8314696123025452370787884; };
template <> struct Root32<22> { static constexpr double re = -0.3826834323650897717284600; static constexpr double im = 0.9238795325112867561281832; };
template <> struct Root32<23> { static constexpr double re = -0.1950903220161282678482849; static constexpr double im = 0.9807852804032304491261822; };
template <> struct Root32<24> { static constexpr double re = 0.0; static constexpr double im = 1.0; };
template <> struct Root32<25> { static constexpr double re = 0.1950903220161282678482849; static constexpr double im = 0.9807852804032304491261822; };
template <> struct Root32<26> { static constexpr double re = 0.3826834323650897717284600; static constexpr double im = 0.9238795325112867561281832; };
template <> struct Root32<27> { static constexpr double re = 0.5555702330196022247428308; static constexpr double im = 0.8314696123025452370787884; };
template <> struct Root32<28> { static constexpr double re = 0.7071067811865475244008444; static constexpr double im = 0.7071067811865475244008444; };
template <> struct Root32<29> { static constexpr double re = 0.8314696123025452370787884; static constexpr double im = 0.5555702330196022247428308; };
template <> struct Root32<30> { static constexpr double re = 0.9238795325112867561281832; static constexpr double im = 0.3826834323650897717284600; };
template <> struct Root32<31> { static constexpr double re = 0.9807852804032304491261822; static constexpr double im = 0.1950903220161282678482849; };
}  // namespace offtb
