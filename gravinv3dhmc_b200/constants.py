"""Physical constants and unit conversions (same names and values as the reference's
constants.py:29-44)."""
SI2MGAL = 100000.0            # constants.py:29
G_SI = 0.00000000006673       # constants.py:33 (m^3 kg^-1 s^-1; the reference calls it G_SPHERICAL)
G = 0.00000006673             # constants.py:34 (density in g/cm^3)
MEAN_EARTH_RADIUS = 6378137.0  # constants.py:44
SI2EOTVOS = 1000000000.0      # constants.py:26
CM = 10. ** (-7)              # constants.py:37
T2NT = 10. ** (6)             # constants.py:41
g0 = 9.80                     # constants.py:50
