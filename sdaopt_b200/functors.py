"""Registered device objectives and the resolution of the reference's ``func`` argument.

The reference takes any Python callable ``f(x, *args)`` (_sda.py:440-444).  The B200 path
evaluates objectives inside the annealing kernel, so ``func`` must name a device functor:

* a :class:`DeviceFunctor` instance (``Rastrigin()``, ``ShiftedRastrigin(3.14159)`` ...),
* the bound method ``Benchmark.fun`` of a catalogue class whose name is registered
  (sdaopt/benchmark/go_benchmark_functions), as the suite passes it (benchmark/bench.py:298-299),
* the function ``sutton_chen`` (sdaopt/benchmark/suttonchen.py:23), matched by name,
* or the registered name as a string.

Anything else raises ``TypeError``: there is no CPU fallback.
"""
import ctypes as C
import hashlib
import os

import numpy as np

from . import _lib

# name -> id, identical to include/sdaopt_b200.h (SDA_FN_*)
FUNCTOR_IDS = dict(Rastrigin=0, ShiftedRastrigin=1, Rosenbrock=2, Ackley01=3, Schwefel01=4,
                   Schwefel26=5, Exponential=6, SuttonChen=7, Constant=8, Sphere=9, Griewank=10)

FUNCTOR_IDS.update({
    'Alpine01': 11,
    'Brown': 12,
    'Cigar': 13,
    'CosineMixture': 14,
    'Deb01': 15,
    'DixonPrice': 16,
    'Easom': 17,
    'Qing': 18,
    'Quintic': 19,
    'Salomon': 20,
    'Schwefel20': 21,
    'Schwefel21': 22,
    'Schwefel22': 23,
    'Step': 24,
    'StyblinskiTang': 25,
    'Trid': 26,
    'Zacharov': 27,
    'SixHumpCamel': 28,
    'Beale': 29,
    'HimmelBlau': 30,
    'Matyas': 31,
    'McCormick': 32,
    'GoldsteinPrice': 33,
    'Levy13': 34,
    'Bohachevsky1': 35,
    'Bukin06': 36,
    'EggCrate': 37,
    'Schaffer01': 38,
    'DropWave': 39,
    'ThreeHumpCamel': 40,
    'Zettl': 41,
    'Leon': 42,
    'Cube': 43,
    'Brent': 44,
    'Price01': 45,
    'Bird': 46,
    'Ackley02': 47,
    'Ackley03': 48,
    'Zirilli': 49,
    'Wavy': 50,
    'Step2': 51,
    'Plateau': 52,
    'Sodp': 53,
    'Trigonometric02': 54,
    'XinSheYang02': 55,
    'SineEnvelope': 56,
    'Pathological': 57,
    'Rana': 58,
    'Adjiman': 59,
})

# catalogue part 3 (ids 60..): the remaining go_funcs_*.py classes, device code in
# csrc/sda_objectives_cat.cuh; metadata introspected from the reference classes
FUNCTOR_IDS.update({
    'AMGM': 60,
    'Alpine02': 61,
    'BartelsConn': 62,
    'BiggsExp02': 63,
    'BiggsExp03': 64,
    'BiggsExp04': 65,
    'BiggsExp05': 66,
    'Bohachevsky2': 67,
    'Bohachevsky3': 68,
    'BoxBetts': 69,
    'Branin01': 70,
    'Branin02': 71,
    'Bukin02': 72,
    'Bukin04': 73,
    'CarromTable': 74,
    'Chichinadze': 75,
    'Cola': 76,
    'Colville': 77,
    'Corana': 78,
    'CrossInTray': 79,
    'CrossLegTable': 80,
    'CrownedCross': 81,
    'Csendes': 82,
    'Damavandi': 83,
    'DeVilliersGlasser01': 84,
    'DeVilliersGlasser02': 85,
    'Deb03': 86,
    'Decanomial': 87,
    'Deceptive': 88,
    'DeckkersAarts': 89,
    'DeflectedCorrugatedSpring': 90,
    'Dolan': 91,
    'Eckerle4': 92,
    'EggHolder': 93,
    'ElAttarVidyasagarDutta': 94,
    'Exp2': 95,
    'FreudensteinRoth': 96,
    'Gear': 97,
    'Giunta': 98,
    'Gulf': 99,
    'Hansen': 100,
    'Hartmann3': 101,
    'Hartmann6': 102,
    'HelicalValley': 103,
    'HolderTable': 104,
    'Hosaki': 105,
    'Infinity': 106,
    'JennrichSampson': 107,
    'Judge': 108,
    'Katsuura': 109,
    'Keane': 110,
    'Kowalik': 111,
    'Langermann': 112,
    'LennardJones': 113,
    'Levy03': 114,
    'Levy05': 115,
    'Meyer': 116,
    'Michalewicz': 117,
    'MieleCantrell': 118,
    'Mishra01': 119,
    'Mishra02': 120,
    'Mishra03': 121,
    'Mishra04': 122,
    'Mishra05': 123,
    'Mishra06': 124,
    'Mishra07': 125,
    'Mishra08': 126,
    'Mishra09': 127,
    'Mishra10': 128,
    'Mishra11': 129,
    'MultiModal': 130,
    'NeedleEye': 131,
    'NewFunction01': 132,
    'NewFunction02': 133,
    'OddSquare': 134,
    'Parsopoulos': 135,
    'Paviani': 136,
    'PenHolder': 137,
    'Penalty01': 138,
    'Penalty02': 139,
    'PermFunction01': 140,
    'PermFunction02': 141,
    'Pinter': 142,
    'Powell': 143,
    'PowerSum': 144,
    'Price02': 145,
    'Price03': 146,
    'Price04': 147,
    'Quadratic': 148,
    'Ratkowsky01': 149,
    'Ratkowsky02': 150,
    'Ripple01': 151,
    'Ripple25': 152,
    'RosenbrockModified': 153,
    'RotatedEllipse01': 154,
    'RotatedEllipse02': 155,
    'Sargan': 156,
    'Schaffer02': 157,
    'Schaffer03': 158,
    'Schaffer04': 159,
    'Schwefel02': 160,
    'Schwefel04': 161,
    'Schwefel06': 162,
    'Schwefel36': 163,
    'Shekel05': 164,
    'Shekel07': 165,
    'Shekel10': 166,
    'Shubert01': 167,
    'Shubert03': 168,
    'Shubert04': 169,
    'StretchedV': 170,
    'TestTubeHolder': 171,
    'Thurber': 172,
    'Treccani': 173,
    'Trefethen': 174,
    'Trigonometric01': 175,
    'Tripod': 176,
    'Ursem01': 177,
    'Ursem03': 178,
    'Ursem04': 179,
    'UrsemWaves': 180,
    'VenterSobiezcczanskiSobieski': 181,
    'Vincent': 182,
    'Watson': 183,
    'WayburnSeader01': 184,
    'WayburnSeader02': 185,
    'Weierstrass': 186,
    'Whitley': 187,
    'Wolfe': 188,
    'XinSheYang03': 189,
    'XinSheYang04': 190,
    'Xor': 191,
    'YaoLiu04': 192,
    'YaoLiu09': 193,
    'ZeroSum': 194,
    'Zimmerman': 195,
    'Stochastic': 196,
    'XinSheYang01': 197,
})

# LennardJones.minima (go_funcs_L.py): best known energies for 2..20 atoms
_LJ_MINIMA = [-1.0, -3.0, -6.0, -9.103852, -12.712062, -16.505384, -19.821489, -24.113360,
              -28.422532, -32.765970, -37.967600, -44.326801, -47.845157, -52.322627, -56.815742,
              -61.317995, -66.530949, -72.659782, -77.1777043]

_CATALOGUE3 = {
    'AMGM': (2, True, (0.0, 10.0), 0.0),
    'Alpine02': (2, True, (0.0, 10.0), -6.1295),
    'BartelsConn': (2, False, (-500.0, 500.0), 1.0),
    'BiggsExp02': (2, False, (0.0, 20.0), 0.0),
    'BiggsExp03': (3, False, (0.0, 20.0), 0.0),
    'BiggsExp04': (4, False, (0.0, 20.0), 0.0),
    'BiggsExp05': (5, False, (0.0, 20.0), 0.0),
    'Bohachevsky2': (2, False, (-100.0, 100.0), 0.0),
    'Bohachevsky3': (2, False, (-100.0, 100.0), 0.0),
    'BoxBetts': (3, False, [(0.9, 1.2), (9.0, 11.2), (0.9, 1.2)], 0.0),
    'Branin01': (2, False, [(-5.0, 10.0), (0.0, 15.0)], 0.39788735772973816),
    'Branin02': (2, False, (-5.0, 15.0), 5.558914403893825),
    'Bukin02': (2, False, [(-15.0, -5.0), (-3.0, 3.0)], -124.75),
    'Bukin04': (2, False, [(-15.0, -5.0), (-3.0, 3.0)], 0.0),
    'CarromTable': (2, False, (-10.0, 10.0), -24.15681551650653),
    'Chichinadze': (2, False, (-30.0, 30.0), -42.94438701899098),
    'Cola': (17, False, [(0.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0), (-4.0, 4.0)], 11.7464),
    'Colville': (4, False, (-10.0, 10.0), 0.0),
    'Corana': (4, False, (-5.0, 5.0), 0.0),
    'CrossInTray': (2, False, (-10.0, 10.0), -2.062611870822739),
    'CrossLegTable': (2, False, (-10.0, 10.0), -1.0),
    'CrownedCross': (2, False, (-10.0, 10.0), 0.0001),
    'Csendes': (2, True, (-1.0, 1.0), float("nan")),
    'Damavandi': (2, False, (0.0, 14.0), float("nan")),
    'DeVilliersGlasser01': (4, False, (1.0, 100.0), 0.0),
    'DeVilliersGlasser02': (5, False, (1.0, 60.0), 0.0),
    'Deb03': (2, True, (-1.0, 1.0), -1.0),
    'Decanomial': (2, False, (-10.0, 10.0), 0.0),
    'Deceptive': (2, True, (0.0, 1.0), -1.0),
    'DeckkersAarts': (2, False, (-20.0, 20.0), -24776.518342168),
    'DeflectedCorrugatedSpring': (2, True, (0.0, 10.0), -1.0),
    'Dolan': (5, False, (-100.0, 100.0), 0.0),
    'Eckerle4': (3, False, [(0.0, 20.0), (1.0, 20.0), (10.0, 600.0)], 0.0014635887487),
    'EggHolder': (2, True, (-512.1, 512.0), -959.640662711),
    'ElAttarVidyasagarDutta': (2, False, (-100.0, 100.0), 1.712780354),
    'Exp2': (2, False, (0.0, 20.0), 0.0),
    'FreudensteinRoth': (2, False, (-10.0, 10.0), 0.0),
    'Gear': (4, False, (12.0, 60.0), 2.7e-12),
    'Giunta': (2, False, (-1.0, 1.0), 0.06447042053690566),
    'Gulf': (3, False, (0.0, 50.0), 0.0),
    'Hansen': (2, False, (-10.0, 10.0), -176.54179),
    'Hartmann3': (3, False, (0.0, 1.0), -3.8627821478),
    'Hartmann6': (6, False, (0.0, 1.0), -3.32236801141551),
    'HelicalValley': (3, False, (-10.0, 10.0), 0.0),
    'HolderTable': (2, False, (-10.0, 10.0), -19.20850256788675),
    'Hosaki': (2, False, [(0.0, 5.0), (0.0, 6.0)], -2.3458115),
    'Infinity': (2, True, (-1.0, 1.0), 0.0),
    'JennrichSampson': (2, False, (-1.0, 1.0), 124.3621824),
    'Judge': (2, False, (-10.0, 10.0), 16.0817307),
    'Katsuura': (2, True, (0.0, 100.0), 1.0),
    'Keane': (2, False, (0.0, 10.0), 0.0),
    'Kowalik': (4, False, (-5.0, 5.0), 0.0003074861),
    'Langermann': (2, False, (0.0, 10.0), -5.1621259),
    'LennardJones': (6, True, (-4.0, 4.0), lambda n: _LJ_MINIMA[n // 3 - 2]),
    'Levy03': (2, False, (-10.0, 10.0), 0.0),
    'Levy05': (2, False, (-10.0, 10.0), -176.1375779),
    'Meyer': (3, False, [(0.0, 1.0), (100.0, 1000.0), (100.0, 500.0)], 87.945855171),
    'Michalewicz': (2, False, (0.0, 3.141592653589793), -1.8013),
    'MieleCantrell': (4, False, (-1.0, 1.0), 0.0),
    'Mishra01': (2, True, (0.0, 1.000000001), 2.0),
    'Mishra02': (2, True, (0.0, 1.000000001), 2.0),
    'Mishra03': (2, False, (-10.0, 10.0), -0.19990562),
    'Mishra04': (2, False, (-10.0, 10.0), -0.177715264826),
    'Mishra05': (2, False, (-10.0, 10.0), -1.019829519930646),
    'Mishra06': (2, False, (-10.0, 10.0), -2.28395),
    'Mishra07': (2, True, (-10.0, 10.0), 0.0),
    'Mishra08': (2, False, (-10.0, 10.0), 0.0),
    'Mishra09': (3, False, (-10.0, 10.0), 0.0),
    'Mishra10': (2, False, (-10.0, 10.0), 0.0),
    'Mishra11': (2, True, (-10.0, 10.0), 0.0),
    'MultiModal': (2, True, (-10.0, 10.0), 0.0),
    'NeedleEye': (2, True, (-10.0, 10.0), 1.0),
    'NewFunction01': (2, False, (-10.0, 10.0), -0.184648852475),
    'NewFunction02': (2, False, (-10.0, 10.0), -0.199409030092),
    'OddSquare': (2, False, (-15.707963267948966, 15.707963267948966), -1.00846728102),
    'Parsopoulos': (2, False, (-5.0, 5.0), 0.0),
    'Paviani': (10, False, (2.001, 9.999), -45.7784684040686),
    'PenHolder': (2, False, (-11.0, 11.0), -0.9635348327265058),
    'Penalty01': (2, True, (-50.0, 50.0), 0.0),
    'Penalty02': (2, True, (-50.0, 50.0), 0.0),
    'PermFunction01': (2, True, lambda n: (-float(n), n + 1.0), 0.0),
    'PermFunction02': (2, True, lambda n: (-float(n), n + 1.0), 0.0),
    'Pinter': (2, True, (-10.0, 10.0), 0.0),
    'Powell': (4, False, (-4.0, 5.0), 0.0),
    'PowerSum': (4, False, (0.0, 4.0), 0.0),
    'Price02': (2, False, (-10.0, 10.0), 0.9),
    'Price03': (2, False, (-5.0, 5.0), 0.0),
    'Price04': (2, False, (-50.0, 50.0), 0.0),
    'Quadratic': (2, True, (-10.0, 10.0), -3873.72418),
    'Ratkowsky01': (4, False, [(0.0, 1000.0), (1.0, 20.0), (0.0, 3.0), (0.1, 6.0)], 8786.404908),
    'Ratkowsky02': (3, False, [(10.0, 200.0), (0.5, 5.0), (0.01, 0.5)], 8.0565229338),
    'Ripple01': (2, False, (0.0, 1.0), -2.2),
    'Ripple25': (2, False, (0.0, 1.0), -2.0),
    'RosenbrockModified': (2, False, (-2.0, 2.0), 34.040243106640844),
    'RotatedEllipse01': (2, False, (-500.0, 500.0), 0.0),
    'RotatedEllipse02': (2, False, (-500.0, 500.0), 0.0),
    'Sargan': (2, True, (-100.0, 100.0), 0.0),
    'Schaffer02': (2, False, (-100.0, 100.0), 0.0),
    'Schaffer03': (2, False, (-100.0, 100.0), 0.00156685),
    'Schaffer04': (2, False, (-100.0, 100.0), 0.292579),
    'Schwefel02': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel04': (2, True, (0.0, 10.0), 0.0),
    'Schwefel06': (2, False, (-100.0, 100.0), 0.0),
    'Schwefel36': (2, False, (0.0, 500.0), -3456.0),
    'Shekel05': (4, False, (0.0, 10.0), -10.1531996791),
    'Shekel07': (4, False, (0.0, 10.0), -10.4029405668),
    'Shekel10': (4, False, (0.0, 10.0), -10.536409816692023),
    'Shubert01': (2, True, (-10.0, 10.0), -186.7309),
    'Shubert03': (2, True, (-10.0, 10.0), -24.062499),
    'Shubert04': (2, True, (-10.0, 10.0), -29.016015),
    'StretchedV': (2, True, (-10.0, 10.0), 0.0),
    'TestTubeHolder': (2, False, (-10.0, 10.0), -10.872299901558),
    'Thurber': (7, False, [(500.0, 2000.0), (500.0, 2000.0), (100.0, 1000.0), (10.0, 150.0), (0.1, 2.0), (0.1, 1.0), (0.0, 0.2)], 5642.7082397),
    'Treccani': (2, False, (-5.0, 5.0), 0.0),
    'Trefethen': (2, False, (-10.0, 10.0), -3.3068686474),
    'Trigonometric01': (2, True, (0.0, 3.141592653589793), 0.0),
    'Tripod': (2, False, (-100.0, 100.0), 0.0),
    'Ursem01': (2, False, [(-2.5, 3.0), (-2.0, 2.0)], -4.81681406371),
    'Ursem03': (2, False, [(-2.0, 2.0), (-1.5, 1.5)], -3.0),
    'Ursem04': (2, False, (-2.0, 2.0), -1.5),
    'UrsemWaves': (2, False, [(-0.9, 1.2), (-1.2, 1.2)], -8.5536),
    'VenterSobiezcczanskiSobieski': (2, False, (-50.0, 50.0), -400.0),
    'Vincent': (2, True, (0.25, 10.0), lambda n: -float(n)),
    'Watson': (6, False, (-5.0, 5.0), 0.002288),
    'WayburnSeader01': (2, False, (-5.0, 5.0), 0.0),
    'WayburnSeader02': (2, False, (-500.0, 500.0), 0.0),
    'Weierstrass': (2, True, (-0.5, 0.5), 0.0),
    'Whitley': (2, True, (-10.24, 10.24), 0.0),
    'Wolfe': (3, False, (0.0, 2.0), 0.0),
    'XinSheYang03': (2, True, (-20.0, 20.0), -1.0),
    'XinSheYang04': (2, True, (-10.0, 10.0), -1.0),
    'Xor': (9, False, (-1.0, 1.0), 0.9597588),
    'YaoLiu04': (2, True, (-10.0, 10.0), 0.0),
    'YaoLiu09': (2, True, (-5.12, 5.12), 0.0),
    'ZeroSum': (2, True, (-10.0, 10.0), 0.0),
    'Zimmerman': (2, False, (0.0, 100.0), 0.0),
    # fun() draws U[0,1) factors at every call (go_funcs_S.py:1274-1280, go_funcs_X.py:47-51);
    # on the device the factors are Philox noise addressed by the point (sda_objectives_cat.cuh)
    'Stochastic': (2, True, (-5.0, 5.0), 0.0),
    'XinSheYang01': (2, True, (-5.0, 5.0), 0.0),
}

# name: (default N, n-D capable, box or per-dimension bounds, fglob or fglob(N)) -- the class
# attributes of go_benchmark_functions/go_funcs_*.py (`_bounds`, `fglob`, `change_dimensionality`)
_CATALOGUE2 = {
    'Alpine01': (2, True, (-10.0, 10.0), 0.0),
    'Brown': (2, True, (-1.0, 4.0), 0.0),
    'Cigar': (2, True, (-100.0, 100.0), 0.0),
    'CosineMixture': (2, True, (-1.0, 1.0), lambda n: -0.9 * n),
    'Deb01': (2, True, (-1.0, 1.0), -1.0),
    'DixonPrice': (2, True, (-10.0, 10.0), 0.0),
    'Easom': (2, False, (-100.0, 100.0), -1.0),
    'Qing': (2, True, (-500.0, 500.0), 0.0),
    'Quintic': (2, True, (-10.0, 10.0), 0.0),
    'Salomon': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel20': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel21': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel22': (2, True, (-100.0, 100.0), 0.0),
    'Step': (2, True, (-100.0, 100.0), 0.0),
    'StyblinskiTang': (2, True, (-5.0, 5.0), lambda n: -39.16616570377142 * n),
    'Trid': (6, True, (-20.0, 20.0), -50.0),
    'Zacharov': (2, True, (-5.0, 10.0), 0.0),
    'SixHumpCamel': (2, False, (-5.0, 5.0), -1.031628),
    'Beale': (2, False, (-4.5, 4.5), 0.0),
    'HimmelBlau': (2, False, (-5.0, 5.0), 0.0),
    'Matyas': (2, False, (-10.0, 10.0), 0.0),
    'McCormick': (2, False, [(-1.5, 4.0), (-3.0, 3.0)], -1.913222954981037),
    'GoldsteinPrice': (2, False, (-2.0, 2.0), 3.0),
    'Levy13': (2, False, (-10.0, 10.0), 0.0),
    'Bohachevsky1': (2, False, (-100.0, 100.0), 0.0),
    'Bukin06': (2, False, [(-15.0, -5.0), (-3.0, 3.0)], 0.0),
    'EggCrate': (2, False, (-5.0, 5.0), 0.0),
    'Schaffer01': (2, False, (-100.0, 100.0), 0.0),
    'DropWave': (2, False, (-5.12, 5.12), -1.0),
    'ThreeHumpCamel': (2, False, (-5.0, 5.0), 0.0),
    'Zettl': (2, False, (-5.0, 10.0), -0.003791237220468656),
    'Leon': (2, False, (-1.2, 1.2), 0.0),
    'Cube': (2, False, (-10.0, 10.0), 0.0),
    'Brent': (2, False, (-10.0, 10.0), 0.0),
    'Price01': (2, False, (-500.0, 500.0), 0.0),
    'Bird': (2, False, (-6.283185307179586, 6.283185307179586), -106.7645367198034),
    'Ackley02': (2, False, (-32.0, 32.0), -200.0),
    'Ackley03': (2, False, (-32.0, 32.0), -195.6290282592388),
    'Zirilli': (2, False, (-10.0, 10.0), -0.35238603),
    'Wavy': (2, True, (-3.141592653589793, 3.141592653589793), 0.0),
    'Step2': (2, True, (-100.0, 100.0), 0.5),
    'Plateau': (2, True, (-5.12, 5.12), 30.0),
    'Sodp': (2, True, (-1.0, 1.0), 0.0),
    'Trigonometric02': (2, True, (-500.0, 500.0), 1.0),
    'XinSheYang02': (2, True, (-6.283185307179586, 6.283185307179586), 0.0),
    'SineEnvelope': (2, True, (-100.0, 100.0), 0.0),
    'Pathological': (2, False, (-100.0, 100.0), 0.0),
    'Rana': (2, True, (-500.000001, 500.000001), -500.8021602966615),
    'Adjiman': (2, False, [(-1.0, 2.0), (-1.0, 1.0)], -2.02180678),
}

_CATALOGUE_ALL = dict(_CATALOGUE2)
_CATALOGUE_ALL.update(_CATALOGUE3)

# default box and known optimum of the catalogue classes (go_funcs_*.py `__init__`)
_CATALOGUE = {
    "Rastrigin": ((-5.12, 5.12), 0.0),      # go_funcs_R.py:80-86
    "Rosenbrock": ((-30.0, 30.0), 0.0),     # go_funcs_R.py:280-286
    "Ackley01": ((-35.0, 35.0), 0.0),       # go_funcs_A.py:38-41
    "Schwefel01": ((-100.0, 100.0), 0.0),   # go_funcs_S.py:324-330
    "Schwefel26": ((-500.0, 500.0), 0.0),   # go_funcs_S.py:606-611
    "Exponential": ((-1.0, 1.0), -1.0),     # go_funcs_E.py:296-301
    "Sphere": ((-5.12, 5.12), 0.0),         # go_funcs_S.py:1149-1154
    "Griewank": ((-100.0, 100.0), 0.0),     # go_funcs_G.py:160-169
}


class DeviceFunctor(object):
    """A registered CUDA objective.  Calling it evaluates on the GPU (one point or a batch)."""

    def __init__(self, name, params=(), dimensions=None):
        if name not in FUNCTOR_IDS:
            raise KeyError("unknown device functor %r" % (name,))
        self.name = name
        self.functor_id = FUNCTOR_IDS[name]
        self.params = np.ascontiguousarray(params, dtype=np.float64).ravel()
        self.N = dimensions
        box, fglob = _CATALOGUE.get(name, (None, float("nan")))
        if name in _CATALOGUE_ALL:
            n_default, _, box, fglob = _CATALOGUE_ALL[name]
            if self.N is None:
                self.N = n_default
            if callable(fglob):
                fglob = fglob(self.N)
            if callable(box):                              # box depends on N (PermFunction01/02)
                box = box(self.N)
        self.fglob = fglob
        self._box = box
        self.library_path = None        # set for user functors: the build that contains them

    @property
    def bounds(self):
        if self._box is None or self.N is None:
            raise AttributeError("%s has no default bounds" % self.name)
        if isinstance(self._box, list):                    # per-dimension bounds (fixed-N classes)
            return list(self._box)
        return [self._box] * int(self.N)

    def fun(self, x, *args):
        return self(x)

    def __call__(self, x, *args):
        x = np.ascontiguousarray(x, dtype=np.float64)
        single = x.ndim == 1
        pts = x.reshape(1, -1) if single else x
        out = evaluate(self, pts)
        return float(out[0]) if single else out

    def __repr__(self):
        return "DeviceFunctor(%s%s)" % (self.name, "" if not self.params.size else
                                        ", params=%s" % self.params.tolist())


def Rastrigin(dimensions=2): return DeviceFunctor("Rastrigin", dimensions=dimensions)
def Rosenbrock(dimensions=2): return DeviceFunctor("Rosenbrock", dimensions=dimensions)
def Ackley01(dimensions=2): return DeviceFunctor("Ackley01", dimensions=dimensions)
def Schwefel01(dimensions=2): return DeviceFunctor("Schwefel01", dimensions=dimensions)
def Schwefel26(dimensions=2): return DeviceFunctor("Schwefel26", dimensions=dimensions)
def Exponential(dimensions=2): return DeviceFunctor("Exponential", dimensions=dimensions)
def Sphere(dimensions=2): return DeviceFunctor("Sphere", dimensions=dimensions)
def Griewank(dimensions=2): return DeviceFunctor("Griewank", dimensions=dimensions)


def ShiftedRastrigin(shift=3.14159, dimensions=None):
    """README.md:61-71: sum((x-s)^2 - 10 cos(2 pi (x-s))) + 10 D"""
    return DeviceFunctor("ShiftedRastrigin", params=[shift], dimensions=dimensions)


def SuttonChen(n_atoms=None):
    """sdaopt/benchmark/suttonchen.py:23-40 (x is the flattened (n_atoms, 3) coordinate array)"""
    return DeviceFunctor("SuttonChen", dimensions=None if n_atoms is None else 3 * n_atoms)


def Constant(value):
    """f(x) = value; the `weirdfunc` of sdaopt/tests/test_sda.py:34"""
    return DeviceFunctor("Constant", params=[value])


def compile_user_functor(source, params=(), name="user", bounds=None, fglob=float("nan"),
                          build_dir=None):
    """User-supplied CUDA device functor (replaces the reference's arbitrary Python ``func``,
    _sda.py:440-444).  ``source`` must define

        template <class X>
        __device__ double sda_user_objective(const X &x, int dim, const double *params);

    where ``x[k]`` is the k-th coordinate.  The annealing / local-search kernels are rebuilt with
    the functor inlined (nvcc, about a minute, cached by source hash under build/user/) and the
    returned DeviceFunctor routes every call to that library.  No CPU evaluation ever happens."""
    from . import build as _build
    # cache key: the functor AND the kernels it is compiled into (a cached build of older kernels
    # must not be picked up after the package's sources changed)
    h = hashlib.sha1(source.encode())
    for dep in sorted(_build.SOURCES + _build.HEADERS):
        path = os.path.join(_build.CSRC, dep)
        if os.path.exists(path):
            with open(path, "rb") as fh:
                h.update(fh.read())
    digest = h.hexdigest()[:16]
    root = build_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                     "build", "user", digest)
    os.makedirs(root, exist_ok=True)
    header = os.path.join(root, "user_functor.cuh")
    so = os.path.join(root, "libsdaopt_b200_user.so")
    if not os.path.exists(so):
        with open(header, "w") as fh:
            fh.write("// user functor %s\n#pragma once\n%s\n" % (name, source))
        _build.build_user(header, so)
    f = DeviceFunctor.__new__(DeviceFunctor)
    f.name = name
    f.functor_id = _lib.SDA_FN_USER
    f.params = np.ascontiguousarray(params, dtype=np.float64).ravel()
    f.N = None if bounds is None else len(bounds)
    f.fglob = fglob
    f._box = None if bounds is None else [tuple(b) for b in bounds]
    f.library_path = so
    return f


def resolve(func):
    """Map the reference's ``func`` argument to a DeviceFunctor or raise TypeError."""
    if isinstance(func, DeviceFunctor):
        return func
    if isinstance(func, str):
        return DeviceFunctor(func)
    owner = getattr(func, "__self__", None)
    if owner is not None and getattr(func, "__name__", "") in ("fun", "_funcwrapped"):
        if isinstance(owner, DeviceFunctor):
            return owner
        # Algo._funcwrapped wraps self._k.fun (benchmark/bench.py:298)
        bench = getattr(owner, "_k", owner)
        name = type(bench).__name__
        if name in FUNCTOR_IDS:
            return DeviceFunctor(name, dimensions=getattr(bench, "N", None))
    if getattr(func, "__name__", "") == "sutton_chen":
        return DeviceFunctor("SuttonChen")
    raise TypeError(
        "sdaopt_b200 evaluates objectives on the GPU: `func` must be a registered device functor "
        "(sdaopt_b200.functors), a catalogue Benchmark.fun bound method, `sutton_chen` or a "
        "registered name -- arbitrary Python callables are not supported (no CPU fallback). "
        "Got %r" % (func,))


def make_problem(functor, lower, upper):
    """(SdaProblem with HOST pointers, keep-alive tuple)"""
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    params = functor.params
    meta = _CATALOGUE_ALL.get(functor.name)
    if meta is not None and not meta[1] and lower.size != meta[0]:
        # fixed-arity class: its fun() indexes x[0..N-1] (the reference raises IndexError on a short x)
        raise ValueError("%s is defined for exactly %d dimensions, got bounds of length %d"
                         % (functor.name, meta[0], lower.size))
    p = _lib.SdaProblem(functor.functor_id, int(lower.size), lower.ctypes.data, upper.ctypes.data,
                        params.ctypes.data if params.size else None, int(params.size))
    return p, (lower, upper, params)


def evaluate(functor, points, device=0):
    """Objective values at points[n, D] (replaces Benchmark.fun / sutton_chen per point)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n, dim = pts.shape
    dummy = np.zeros(dim)
    p, keep = make_problem(functor, dummy, dummy + 1)
    out = np.empty(n, dtype=np.float64)
    _lib.require_cuda()
    _lib.check(_lib.lib(functor.library_path).sda_eval_host(C.byref(p), pts.ctypes.data, n,
                                                         out.ctypes.data, device))
    return out
