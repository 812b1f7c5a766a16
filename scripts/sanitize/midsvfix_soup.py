import ctypes, random, numpy as np, sys
lib=ctypes.CDLL("/tmp/libfix_asan.so")
vp,i64,ci,cd=ctypes.c_void_p,ctypes.c_int64,ctypes.c_int,ctypes.c_double
lib.dj_midsv_postfix.argtypes=[vp,vp,vp,i64,vp,i64,ctypes.c_int32,cd,vp,vp,vp,ci]; lib.dj_midsv_postfix.restype=i64
rnd=random.Random(9)
alphabet=["=A","=C","=N","=n","-A","-C","-c","--C","+A|=T","+A|+C|=T","+A|+C|-G","++A|=T","+|=A","+A","","|","+A|","+A|=N","+c|=n","*AG","-","+","+-A|=C","N","-N","+A|+A|+A|=A","-a","+a|=T","||","+||","+A||=T","|=N"]
for it in range(300):
    rows=[",".join(rnd.choice(alphabet) for _ in range(rnd.randint(0,20))) for _ in range(rnd.randint(1,200))]
    seqlen=rnd.randint(0,25)
    seq=np.frombuffer(bytes(rnd.choice(b"ACGT-") for _ in range(seqlen)) or b"\0",dtype=np.uint8).copy()
    # exact-size buffers so that ASAN sees any overrun
    lens=np.array([len(r) for r in rows],dtype=np.int32)
    off=np.zeros(len(rows),dtype=np.int64); off[1:]=np.cumsum(lens[:-1].astype(np.int64)+1)
    total=int(off[-1]+lens[-1])
    text=(ctypes.c_uint8*max(total,1)).from_buffer_copy(("\n".join(rows)).encode().ljust(max(total,1),b"\n"))
    out=(ctypes.c_uint8*max(total,1))()
    out_len=np.empty(len(rows),dtype=np.int32); keep=np.empty(len(rows),dtype=np.uint8)
    seqbuf=(ctypes.c_uint8*max(seqlen,1)).from_buffer_copy(seq.tobytes()[:max(seqlen,1)].ljust(max(seqlen,1),b"A"))
    for mode in (1,2,4,7,5,3,6):
        rc=lib.dj_midsv_postfix(ctypes.addressof(text),off.ctypes.data,lens.ctypes.data,len(rows),ctypes.addressof(seqbuf),seqlen,mode,95.0,ctypes.addressof(out),out_len.ctypes.data,keep.ctypes.data,rnd.randint(1,4))
        assert rc==0
        assert (out_len<=lens).all()
print("asan run complete")
