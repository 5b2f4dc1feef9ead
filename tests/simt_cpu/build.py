"""TEST INFRASTRUCTURE: builds a CPU-executable copy of SIMT translation units of libcolddiff (see shim/cuda_runtime.h).

`build(['model2_bwd.cu', ...])` rewrites the few CUDA-only constructs textually --
    kernel<<<grid, block, smem, stream>>>(args)   ->  simt::launch(grid, block, smem, [&] { kernel(args); })
    extern __shared__ T name[];                   ->  T* name = (T*)simt::dyn_smem();
    asm [volatile](...);                          ->  simt::unsupported("inline PTX");
-- and compiles the result with g++ together with the fiber scheduler into tests/simt_cpu/_build/lib<tag>.so, which exports
the same extern "C" entry points as the CUDA library for those files.  Nothing here is shipped or imported by the package."""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'cold_diffusion_models_b200', 'csrc')
OUT = os.path.join(HERE, '_build')


def _match_forward(s, i, open_ch, close_ch):
    """s[i] == open_ch -> index just past the matching close_ch (string / char literals skipped)"""
    depth, j = 0, i
    while j < len(s):
        c = s[j]
        if c == '"' or c == "'":
            q = c
            j += 1
            while s[j] != q:
                j += 2 if s[j] == '\\' else 1
        elif c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
        j += 1
    raise ValueError('unbalanced %s' % open_ch)


def _split_top(s):
    parts, depth, cur = [], 0, ''
    for c in s:
        if c in '([{<':
            depth += 1
        elif c in ')]}>':
            depth -= 1
        if c == ',' and depth == 0:
            parts.append(cur.strip()); cur = ''
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def _kernel_expr_start(s, end):
    """index where the kernel name (with optional template arguments) that ends at `end` starts"""
    j = end
    while j > 0 and s[j - 1].isspace():
        j -= 1
    if s[j - 1] == '>':
        depth = 0
        while True:
            j -= 1
            if s[j] == '>':
                depth += 1
            elif s[j] == '<':
                depth -= 1
                if depth == 0:
                    break
        while j > 0 and s[j - 1].isspace():
            j -= 1
    while j > 0 and (s[j - 1].isalnum() or s[j - 1] in '_:'):
        j -= 1
    return j


# device helpers of the library whose bodies are inline PTX with a plain-C meaning on the CPU: asynchronous global->shared
# copies complete at once (so commit / wait are no-ops), cvt.rna.tf32 is round-to-nearest-away on the low 13 mantissa bits
_PTX_HELPER_BODIES = {
    'cd_cp_async4': '{ if (valid) memcpy(smem_dst, gsrc, 4); else memset(smem_dst, 0, 4); }',
    'cd_cp_async16': '{ if (valid) memcpy(smem_dst, gsrc, 16); else memset(smem_dst, 0, 16); }',
    # warp-level m16n8k8 MMA (PTX fragment layout) through lane exchanges: D[g][2t+j] += sum_k A[g][k] B[k][2t+j]
    'cd_mma_m16n8k8_tf32': '''{ const int lane_ = threadIdx.x & 31, g_ = lane_ >> 2, t_ = lane_ & 3;
      float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
      for (int k_ = 0; k_ < 8; ++k_) {
        const int kk_ = k_ & 3, hi_ = k_ >> 2;
        const float ag = __shfl_sync(0xffffffffu, hi_ ? a[2] : a[0], g_ * 4 + kk_);
        const float ah = __shfl_sync(0xffffffffu, hi_ ? a[3] : a[1], g_ * 4 + kk_);
        const float b0_ = __shfl_sync(0xffffffffu, hi_ ? b[1] : b[0], (2 * t_) * 4 + kk_);
        const float b1_ = __shfl_sync(0xffffffffu, hi_ ? b[1] : b[0], (2 * t_ + 1) * 4 + kk_);
        c0 += ag * b0_; c1 += ag * b1_; c2 += ah * b0_; c3 += ah * b1_;
      }
      d[0] += c0; d[1] += c1; d[2] += c2; d[3] += c3; }''',
    'cd_round_tf32': '{ unsigned u = __float_as_uint(x); u = (u + 0x1000u) & ~0x1FFFu; return __uint_as_float(u); }',
}


def rewrite(src):
    for name, body in _PTX_HELPER_BODIES.items():
        m = re.search(r'\b%s\s*\([^)]*\)\s*\{' % name, src)
        if m:
            end = _match_forward(src, m.end() - 1, '{', '}')
            src = src[:m.end() - 1] + body + src[end:]
    src = re.sub(r'asm\s+volatile\s*\(\s*"cp\.async\.(commit_group|wait_group\s+\d+|wait_all);"\s*:::\s*"memory"\s*\)', '(void)0', src)
    src = src.replace('"../../include/colddiff.h"', '"%s"' % os.path.join(ROOT, 'include', 'colddiff.h'))
    # dynamic shared memory
    def dyn(m):
        ty = re.sub(r'__align__\s*\(\s*\d+\s*\)', '', m.group(1)).strip()
        return '%s* %s = reinterpret_cast<%s*>(simt::dyn_smem());' % (ty, m.group(2), ty)
    src = re.sub(r'extern\s+__shared__\s+([^;\[\]]+?)\s+(\w+)\s*\[\s*\]\s*;', dyn, src)
    # inline PTX
    out, i = '', 0
    for m in re.finditer(r'\basm\s*(volatile\s*)?\(', src):
        if m.start() < i:
            continue
        end = _match_forward(src, m.end() - 1, '(', ')')
        semi = src.index(';', end)
        out += src[i:m.start()] + 'simt::unsupported("inline PTX")'
        i = semi
    src = out + src[i:]
    # kernel launches
    out, i = '', 0
    while True:
        k = src.find('<<<', i)
        if k < 0:
            break
        start = _kernel_expr_start(src, k)
        cfg_end = src.index('>>>', k)
        cfg = _split_top(src[k + 3:cfg_end])
        while len(cfg) < 3:
            cfg.append('0')
        a0 = src.index('(', cfg_end)
        a1 = _match_forward(src, a0, '(', ')')
        kernel, args = src[start:k].strip(), src[a0:a1]
        out += src[i:start] + 'simt::launch(dim3(%s), dim3(%s), (size_t)(%s), [&] { %s%s; })' % (cfg[0], cfg[1], cfg[2], kernel, args)
        i = a1
    return out + src[i:]


SIMT_UNITS = ['api.cu', 'elementwise.cu', 'backward.cu', 'degrade.cu', 'conv_simt.cu', 'model2_bwd.cu', 'repack.cu', 'linattn_small.cu', 'layernorm_multi.cu', 'final_proj.cu', 'linattn_ctx.cu', 'linattn_bwd.cu', 'snow_gen.cu']


def build_all():
    """every CUDA-core translation unit + the tensor-core stubs: a CPU library with the complete C ABI of include/colddiff.h.
    SIMT_CPU_ASAN=1 builds it with AddressSanitizer (run python under LD_PRELOAD=$(gcc -print-file-name=libasan.so)
    ASAN_OPTIONS=detect_leaks=0): out-of-bounds accesses of the kernels to heap tensors / shared arrays then abort with a report."""
    asan = os.environ.get('SIMT_CPU_ASAN') == '1'
    return build(SIMT_UNITS, tag='colddiff_cpu_asan' if asan else 'colddiff_cpu', extra_sources=[os.path.join(HERE, 'tc_stubs.cpp')],
                 extra_flags=('-fsanitize=address', '-fno-omit-frame-pointer', '-O1') if asan else ())


def build(units, tag=None, extra_flags=(), extra_sources=()):
    os.makedirs(OUT, exist_ok=True)
    if 'api.cu' not in units:
        units = ['api.cu'] + list(units)
    tag = tag or '_'.join(os.path.splitext(os.path.basename(u))[0] for u in units)
    texts = {}
    for name in list(units) + [f for f in os.listdir(CSRC) if f.endswith('.cuh')]:
        # a unit given as an absolute path (test-only kernels next to this file) keeps its base name in the build directory
        texts[os.path.basename(name)] = rewrite(open(os.path.join(CSRC, name)).read())
    runtime = ''.join(open(f).read() for f in [os.path.join(HERE, 'simt.cpp'), os.path.join(HERE, 'shim', 'cuda_runtime.h')] + list(extra_sources))
    digest = hashlib.sha1(('\0'.join(k + v for k, v in sorted(texts.items())) + runtime + ' '.join(extra_flags)).encode()).hexdigest()[:16]
    lib = os.path.join(OUT, 'lib%s_%s.so' % (tag, digest))
    if os.path.exists(lib):
        return lib
    srcs = []
    for name, text in texts.items():
        path = os.path.join(OUT, name if name.endswith('.cuh') else name + '.cpp')
        with open(path, 'w') as f:
            f.write(text)
        if not name.endswith('.cuh'):
            srcs.append(path)
    cmd = ['g++', '-O2', '-g', '-std=c++17', '-fPIC', '-shared', '-w', '-DCD_HOST_ONLY', '-I' + OUT, '-I' + os.path.join(HERE, 'shim'),
           '-I' + HERE, '-I' + os.path.join(ROOT, 'include')] + list(extra_flags) + srcs + list(extra_sources) + [os.path.join(HERE, 'simt.cpp'), '-o', lib]
    subprocess.check_call(cmd)
    for f in os.listdir(OUT):                                   # older builds of the same library
        if f.startswith('lib%s_' % tag) and f.endswith('.so') and os.path.join(OUT, f) != lib and f[len('lib%s_' % tag):-3].isalnum():
            os.remove(os.path.join(OUT, f))
    return lib


if __name__ == '__main__':
    import sys
    print(build(sys.argv[1:]) if sys.argv[1:] else build_all())
