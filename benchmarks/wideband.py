"""Synthetic wideband captures for the channeliser (bench / test input only): 96 channels of 48 kS/s int16 IQ are placed on a
12.5 kHz raster inside one 1.2 MS/s capture by FFT interpolation (x25) and frequency translation, summed and quantised."""
import torch

M, D = 96, 25


def wideband_from_channels(iq48, amp=500.0, chunk=1920 * 25):
    """iq48: int16 CUDA tensor [96][N][2] (N a multiple of `chunk`-compatible block), channel k centred k*12.5 kHz above the
    capture centre (k >= 48: below).  Each channel is scaled so that its amplitude is `amp` LSB in the capture.
    Returns int16 [25*N][2].  Done block-wise in the frequency domain over the whole record (one FFT per channel)."""
    assert iq48.shape[0] == M and iq48.shape[2] == 2
    N = iq48.shape[1]
    dev = iq48.device
    x = torch.complex(iq48[..., 0].float(), iq48[..., 1].float()) * (amp / 16383.0)        # [96][N]
    X = torch.fft.fft(x, dim=1)                                                             # bins: 48 kHz / N each
    W = torch.zeros(N * D, dtype=torch.complex64, device=dev)
    half = N // 2
    for k in range(M):
        # channel k occupies +-24 kHz around k*12.5 kHz; only +-6.25 kHz of it is kept (neighbours overlap otherwise)
        keep = int(N * 6250 / 48000)
        c = (k if k < M // 2 else k - M) * (N * D // M)                                     # centre bin in the wide spectrum
        idx_pos = torch.arange(0, keep, device=dev)
        idx_neg = torch.arange(-keep, 0, device=dev)
        W[(c + idx_pos) % (N * D)] += X[k, idx_pos]
        W[(c + idx_neg) % (N * D)] += X[k, idx_neg % N]
    w = torch.fft.ifft(W) * D
    out = torch.stack([w.real, w.imag], 1).round().clamp(-32768, 32767).to(torch.int16)
    return out.contiguous()
