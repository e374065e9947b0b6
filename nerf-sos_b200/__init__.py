"""nerfsos_b200 -- B200-native drop-in for the NeRF-SOS volumetric-rendering hot path.

Public surface (same names as the reference):
    nerfsos_b200.models.nerf_net.NeRFNet          <-> models/nerf_net.py:20
    nerfsos_b200.models.nerf_mlp.NeRFMLP / MLP    <-> models/nerf_mlp.py:24,132
    nerfsos_b200.utils.image.CorrelationLoss / GeoCorrelationLoss / img2mse / mse2psnr / get_similarity_matrix
                                                   <-> utils/image.py
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/nerfsos.h
(libnerfsos.so, loaded with ctypes by nerfsos_b200._lib).  There is no CPU fallback.
"""
__version__ = "0.1.0"
