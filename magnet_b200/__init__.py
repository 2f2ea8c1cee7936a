"""magnet_b200 — B200-native hot path of MAgNet (jaggbow/magnet): graph construction, the two
message-passing layer flavours and the INR decoder as hand-written sm_100a CUDA kernels behind
the reference's own module API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
