import ctypes, random, json, numpy as np
lib=ctypes.CDLL("/tmp/libingest_asan.so")
vp,i64,ci=ctypes.c_void_p,ctypes.c_int64,ctypes.c_int
lib.dj_jsonl_count_lines.argtypes=[vp,i64,ci]; lib.dj_jsonl_count_lines.restype=i64
lib.dj_jsonl_index.argtypes=[vp,i64,i64,vp,vp,vp,vp,vp,vp,ci]; lib.dj_jsonl_index.restype=i64
lib.dj_gather_fixed.argtypes=[vp,vp,vp,i64,ctypes.c_int32,vp,ci]
rnd=random.Random(4)
pieces_all=['{"QNAME": "q1", "RNAME": "c", "MIDSV": "=A,=C,+T|=G", "STRAND": "+", "PRESET": "map-ont"}','{"QNAME":"q\\"x","MIDSV":"=A"}','{"MIDSV": "=A,=C"','"QNAME": ','{"QNAME": "a", "MIDSV": ','{}','','   ','{"QNAME": "\\\\", "MIDSV": "\\u0041"}','{"MIDSV": "=A", "QNAME": "z", "STRAND": "-"}','{"QNAME": "q", "MIDSV": "=A,=C", "STRAND": "','"','\\','{"QNAME": "q", "MIDSV": "=A"}\r','{"STRAND": "+"}']
pieces=[json.dumps({"QNAME":"r%d"%i,"RNAME":"c","MIDSV":",".join(random.Random(i).choice(["=A","-C","*AG","+T|+T|=G","=N"]) for _ in range(random.Random(i).randint(1,30))),"STRAND":"+-"[i%2],"PRESET":"map-ont"}) for i in range(40)]+[json.dumps({"MIDSV":"=A,=C","STRAND":"-","QNAME":"zz"})]
ok=irr=0
for it in range(3000):
    k=rnd.randint(0,12)
    lines=[rnd.choice(pieces) if rnd.random()<0.8 else rnd.choice(pieces)[:rnd.randint(0,30)] for _ in range(k)]
    data="\n".join(lines).encode()
    if rnd.random()<0.5: data+=b"\n"
    if rnd.random()<0.1: data=bytes(rnd.randrange(256) for _ in range(rnd.randint(0,80)))
    size=len(data)
    # the product gives 64 readable bytes after the file; ASAN sees anything beyond
    buf=(ctypes.c_uint8*(size+64)).from_buffer_copy(data+b"\n"*64)
    thr=rnd.randint(1,4)
    n=lib.dj_jsonl_count_lines(ctypes.addressof(buf),size,thr)
    assert 0<=n<=size+1
    m=max(n,1)
    off=np.empty(m,np.int64); ln=np.empty(m,np.int32); qo=np.empty(m,np.int64); ql=np.empty(m,np.int32); st=np.empty(m,np.uint8); hs=ctypes.c_int32(0)
    r=lib.dj_jsonl_index(ctypes.addressof(buf),size,n,off.ctypes.data,ln.ctypes.data,qo.ctypes.data,ql.ctypes.data,st.ctypes.data,ctypes.byref(hs),thr)
    if r==0 and n>0:
        ok+=1
        assert (off[:n]>=0).all() and (off[:n]+ln[:n]<=size).all() and (ln[:n]>=0).all(), (data, off[:n], ln[:n])
        assert (qo[:n]>=0).all() and (qo[:n]+ql[:n]<=size).all() and (ql[:n]>=0).all()
        # agreement with json on regular files
        recs=[json.loads(l) for l in data.decode("latin-1").splitlines() if l.strip()]
        assert len(recs)==n
        for i,rec in enumerate(recs):
            assert data[off[i]:off[i]+ln[i]].decode("latin-1")==rec["MIDSV"], (data, i)
            assert data[qo[i]:qo[i]+ql[i]].decode("latin-1")==rec["QNAME"]
        w=int(ql[:n].max()) or 1
        out=np.zeros((n,w),np.uint8)
        lib.dj_gather_fixed(ctypes.addressof(buf),qo.ctypes.data,ql.ctypes.data,n,w,out.ctypes.data,thr)
    else: irr+=1
print("regular files",ok,"irregular",irr)
