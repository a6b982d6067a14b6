"""Random / mutated / hostile .obj bytes into tmpt_load_obj, odd sizes into tmpt_write_png, argument lists into tmpt_main -- against a
sanitized build of csrc/host.cpp (tools/host_fuzz/run.sh builds it and passes its path).  usage: fuzz_host.py <seed> <files> <lib>"""
import ctypes as C, sys, os, numpy as np, tempfile
L = C.CDLL(sys.argv[3])
L.tmpt_last_error.restype = C.c_char_p
rng = np.random.default_rng(int(sys.argv[1]))
td = tempfile.mkdtemp()
p = os.path.join(td, "f.obj").encode()
good = b"v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\nf 1/1/1 2/1/1 3/1/1\nf -1 -2 -3\n# c\nusemtl x\ng y\n"
alphabet = b"vfntg 0123456789.-+eE/\n\r\t#usemtl\x00\xff"
n_ok = n_err = 0
for it in range(int(sys.argv[2])):
    k = rng.integers(0, 4)
    if k == 0:   # random bytes from the alphabet
        data = bytes(rng.choice(list(alphabet), rng.integers(0, 400)).astype(np.uint8))
    elif k == 1:  # mutate a good file
        b = bytearray(good * int(rng.integers(1, 6)))
        for _ in range(int(rng.integers(1, 12))):
            i = int(rng.integers(0, len(b))); b[i] = int(rng.choice(list(alphabet)))
        data = bytes(b)
    elif k == 2:  # huge / negative / zero indices
        data = good + b"f %d %d %d\n" % tuple(int(x) for x in rng.integers(-2**33, 2**33, 3)) + b"f 0 1 2\nf 1 2\nf\nf 1 2 3 4 5 6 7 8 9 10\n"
    else:         # very long line, no newline at the end, numbers with many digits
        data = b"v " + b"9" * int(rng.integers(1, 5000)) + b" 1e" + b"9" * int(rng.integers(1, 12)) + b" -." + b"0" * 400 + b"1\n" + good[:-1] + b" " * int(rng.integers(0, 70000))
    open(p, "wb").write(data)
    tris = C.POINTER(C.c_float)(); cnt = C.c_int(0); mn = (C.c_float * 3)(); mx = (C.c_float * 3)()
    rc = L.tmpt_load_obj(p, C.byref(tris), C.byref(cnt), mn, mx)
    if rc == 0:
        n_ok += 1
        a = np.ctypeslib.as_array(tris, shape=(cnt.value * 9,)).copy()  # touch every float the loader says it wrote
        L.tmpt_free(tris)
    else:
        n_err += 1
print("loaded", n_ok, "rejected", n_err)
# PNG writer with odd sizes
for (w, h) in [(1, 1), (3, 2), (16385, 2), (2, 16385)]:
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    assert L.tmpt_write_png(os.path.join(td, "o.png").encode(), w, h, img.ctypes.data_as(C.c_void_p), 1) == 0
# the command line's argument handling
L.tmpt_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
for args in ([b"x"], [b"x", b"1", b"2", b"3"], [b"x", b"0", b"1", b"1", p], [b"x", b"10", b"10", b"1", p], [b"x", b"10", b"10", b"1", b"/nonexistent"]):
    arr = (C.c_char_p * len(args))(*args)
    L.tmpt_main(len(args), arr)
print("host glue clean")
