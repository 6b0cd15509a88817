"""datmo_using_optical_flow_b200 — the B200-native hot path of
anvithaanchala/DATMO_using_Optical_flow's Optical_flow pipeline.

LiDAR cloud -> BEV grid -> dense Farneback flow -> velocity grid -> moving-cell
mask -> DBSCAN labels, as hand-written CUDA for sm_100a behind a C ABI
(include/datmo_b200.h, csrc/).  ``main`` mirrors the reference's function
names; ``engine`` is the device-resident batch interface; ``synth`` generates
the seeded synthetic inputs.  No CPU fallback: compute functions raise when
the CUDA library or a B200 is missing.
"""
__version__ = "0.1.0"
