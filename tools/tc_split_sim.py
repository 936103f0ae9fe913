"""CPU simulation: max |dp| of the v5 model when its GEMMs use split-precision tensor-core arithmetic,
including the FP16 two-way split with per-stream scaling the TC16 build uses (fp16x3; fp16x3_unscaled shows why it scales).
Reads the reference ONNX weights through the oracle; float64 run = ground truth."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import vad_oracle as vo  # noqa: E402

ONNX = str(ROOT / 'cutter-vad_b200/real_time_vad/models/silero_vad_v5.onnx')
W = vo.v5_named_weights(ONNX)
def bf16(x):
    u = x.astype(np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return (((u + r) >> 16) << 16).astype(np.uint32).view(np.float32)
def tf32_trunc(x):
    u = x.astype(np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)
def split(x, f, n):
    parts=[]; r = x.astype(np.float32)
    for i in range(n):
        p = f(r); parts.append(p); r = (r - p).astype(np.float32)
    return parts
def f16(x):
    with np.errstate(over='ignore'):
        return x.astype(np.float32).astype(np.float16).astype(np.float32)
def pow2_at_least(x):
    return 2.0 ** np.ceil(np.log2(np.maximum(x, 1e-30)))
def mm_fp16x3(scaled):
    """FP16 two-way split, 3 products (CVAD_MATH_TC16).  scaled=False: operands as they are (FP16 range problems show);
    scaled=True: every activation row (= stream) and the weight matrix scaled by the power of two that brings its
    largest element to [2^14, 2^15), as the kernels do."""
    def mm(A, Wt):
        A = A.astype(np.float32); Wt = Wt.astype(np.float32)
        if scaled:
            sa = (2.0 ** 15 / pow2_at_least(np.abs(A).max(1, keepdims=True))).astype(np.float32)
            sa = np.where(np.abs(A).max(1, keepdims=True) > 0, sa, 1.0).astype(np.float32)
            sw = np.float32(2.0 ** 15 / pow2_at_least(np.abs(Wt).max()))
        else:
            sa = np.ones((A.shape[0], 1), np.float32); sw = np.float32(1.0)
        As = A * sa; Ws = Wt * sw
        a0 = f16(As); a1 = f16(As - a0); w0 = f16(Ws); w1 = f16(Ws - w0)
        with np.errstate(invalid='ignore', over='ignore'):
            corr = (a1 @ w0.T).astype(np.float32) + (a0 @ w1.T).astype(np.float32)
            acc = (a0 @ w0.T).astype(np.float32) + corr
        return acc / (sa * sw)
    return mm
def mm_factory(mode):
    if mode=='f64': return lambda A,Wt: A.astype(np.float64) @ Wt.astype(np.float64).T
    if mode=='f32': return lambda A,Wt: A.astype(np.float32) @ Wt.astype(np.float32).T
    if mode=='fp16x3_unscaled': return mm_fp16x3(False)
    if mode=='fp16x3': return mm_fp16x3(True)
    f, na, nw, terms = {
      'bf16x1': (bf16,1,1,[(0,0)]),
      'tf32x1': (tf32_trunc,1,1,[(0,0)]),
      'bf16x3': (bf16,2,2,[(0,0),(0,1),(1,0)]),
      'bf16x4': (bf16,2,2,[(0,0),(0,1),(1,0),(1,1)]),
      'tf32x3': (tf32_trunc,2,2,[(0,0),(0,1),(1,0)]),
      'bf16x6': (bf16,3,3,[(0,0),(0,1),(1,0),(0,2),(2,0),(1,1)]),
      'bf16_a2w3x5': (bf16,2,3,[(0,0),(0,1),(1,0),(1,1),(0,2)]),
    }[mode]
    def mm(A,Wt):
        As = split(A,f,na); Ws = split(Wt,f,nw)
        acc = np.zeros((A.shape[0], Wt.shape[0]), np.float32)
        for (i,j) in reversed(terms):
            acc = acc + (As[i] @ Ws[j].T).astype(np.float32)
        return acc
    return mm
def sigmoid(x): return 1/(1+np.exp(-x))
class V5:
    def __init__(s, mode, dt=np.float32):
        s.mm = mm_factory(mode); s.dt = np.float64 if mode=='f64' else np.float32
        s.basis = W['stft.forward_basis_buffer'][:,0,:]
        s.enc = [(W[f'encoder.{i}.reparam_conv.weight'], W[f'encoder.{i}.reparam_conv.bias']) for i in range(4)]
        s.wih=W['decoder.rnn.weight_ih']; s.whh=W['decoder.rnn.weight_hh']; s.b=(W['decoder.rnn.bias_ih'].astype(np.float64)+W['decoder.rnn.bias_hh']).astype(s.dt)
        s.dw=W['decoder.decoder.2.weight'].reshape(128); s.db=W['decoder.decoder.2.bias'].reshape(())
    def conv(s, x, w, b, stride):   # x [B,C,T] pad 1 k3
        B,C,T = x.shape
        xp = np.zeros((B,C,T+2), s.dt); xp[:,:,1:T+1]=x
        To = (T+2-3)//stride+1
        out = np.zeros((B,w.shape[0],To), s.dt)
        for t in range(To):
            col = xp[:,:,t*stride:t*stride+3].reshape(B, C*3)
            out[:,:,t] = s.mm(col, w.reshape(w.shape[0], C*3)) + b
        return np.maximum(out,0)
    def frame(s, x, h, c):   # x [B,512]
        B = x.shape[0]
        cols = [s.mm(x[:, t*128:t*128+256], s.basis) for t in range(3)]
        st = np.stack(cols, axis=2).astype(s.dt)   # [B,258,3]
        mag = np.sqrt(st[:,:129]**2 + st[:,129:]**2).astype(s.dt)
        e = mag
        for i,stride in enumerate([1,2,2,1]):
            e = s.conv(e, s.enc[i][0], s.enc[i][1], stride)
        xf = e[:,:,0]
        g = s.mm(np.concatenate([xf,h],1), np.concatenate([s.wih,s.whh],1)) + s.b
        i_,f_,g_,o_ = g[:,:128], g[:,128:256], g[:,256:384], g[:,384:]
        c = sigmoid(f_)*c + sigmoid(i_)*np.tanh(g_)
        h = sigmoid(o_)*np.tanh(c)
        h = h.astype(s.dt); c = c.astype(s.dt)
        p = sigmoid(np.maximum(h,0) @ s.dw.astype(s.dt) + s.db)
        return p, h, c
def run(mode, audio):   # audio [B, L]
    m = V5(mode); B = audio.shape[0]
    h = np.zeros((B,128), m.dt); c = np.zeros((B,128), m.dt)
    T = audio.shape[1]//512; P = np.zeros((B,T))
    for t in range(T):
        p,h,c = m.frame(audio[:, t*512:(t+1)*512].astype(m.dt), h, c); P[:,t]=p
    return P
if __name__ == '__main__':
    import wave
    w = wave.open(str(ROOT / 'tests/golden/SampleVoiceMono.wav')); raw = np.frombuffer(w.readframes(w.getnframes()), np.int16).astype(np.float32)/32768
    a16 = vo.resample(raw, 48000, 16000)
    rng = np.random.default_rng(0)
    L = 512*200
    streams = [a16[:L], a16[L:2*L], (0.1*rng.standard_normal(L)).astype(np.float32), (0.005*rng.standard_normal(L)).astype(np.float32)]
    t = np.arange(L)/16000
    v = (0.4*np.sin(2*np.pi*150*t)+0.3*np.sin(2*np.pi*300*t)+0.2*np.sin(2*np.pi*600*t))*0.7 + 0.1*rng.standard_normal(L)
    gate = (np.floor(t/0.7)%2==0)
    streams.append((v*gate).astype(np.float32)); streams.append((a16[:L]*0.05).astype(np.float32))
    streams.append(np.where(np.abs(a16[:L])>0.01, a16[:L], 0).astype(np.float32))
    streams.append((a16[:L]*32767).astype(np.float32)); streams.append((a16[:L]*1e-4).astype(np.float32))   # out-of-range amplitudes
    A = np.stack(streams)
    ref = run('f64', A)
    print('ref prob range', ref.min(), ref.max())
    for mode in sys.argv[1:] or ['f32','tf32x1','bf16x1','bf16x3','bf16x4','tf32x3','bf16x6','bf16_a2w3x5','fp16x3_unscaled','fp16x3']:
        P = run(mode, A); d = np.abs(P-ref)
        print(f'{mode:12s} max|dp|={d.max():.3e} per-stream max={np.array2string(d.max(1), precision=1)} mean={d.mean():.2e}')
