"""Physical constants, built with the same expressions (hence the same bits)
as ``ocelot/common/globals.py:13-24``."""
pi = 3.141592653589793
speed_of_light = 299792458.0          # m/s
q_e = 1.6021766208e-19                # C
m_e_kg = 9.10938215e-31               # kg
m_e_eV = m_e_kg * speed_of_light ** 2 / q_e
m_e_MeV = m_e_eV / 1e+6
m_e_GeV = m_e_eV / 1e+9
mu_0 = 4 * pi * 1e-7
epsilon_0 = 1 / mu_0 / speed_of_light ** 2
