#!/usr/bin/env python3
"""Long-running fuzz of the DEVICE code under host emulation (tests/hostemu) against the CPU oracle: a million random
blocks through every decoder, the compressed-domain operations on random block streams, PVRTC up to 256^2 in 1/5/8
stripes, 250,000 random ETC1 blocks per strategy, megapixel DXT images in all four formats under both warp-vote
answers.  Test infrastructure; takes about a minute.  Run tests/test_hostemu.py once first (it builds the library).
Last run (round 2, final kernels, ETC1 with the line and dot forms): 0 mismatches."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import checkers as ck, imagegen
lib=C.CDLL(os.path.join(ROOT, "tests", "hostemu", "libhostemu.so")); u8p=C.POINTER(C.c_uint8)
P=lambda a:a.ctypes.data_as(u8p)
lib.emu_encode4x4.argtypes=[C.c_int,C.c_int,u8p]+[C.c_uint32]*5+[C.c_int,C.c_int,u8p]
lib.emu_pvrtc2.argtypes=[u8p,C.c_uint32,C.c_uint32,C.c_uint32,u8p]
lib.emu_decode4x4.argtypes=[C.c_int,u8p,C.c_uint32,C.c_uint32,C.c_uint32,C.c_int,u8p,C.c_uint32]
lib.emu_downsample4x4.argtypes=[C.c_int,C.c_int,u8p,C.c_uint32,C.c_uint32,u8p]
lib.emu_pad4x4.argtypes=[C.c_int,C.c_int,u8p]+[C.c_uint32]*4+[u8p]
lib.emu_transcode_dxt1_to_etc1.argtypes=[u8p,C.c_uint64]
rng=np.random.default_rng(2026)
t0=time.time(); bad=[]
# decoders: 1M random blocks each
for codec in (0,1,2):
    nc=4 if codec==1 else 3; h,w=2048,2048
    blocks=rng.integers(0,256,(h//4)*(w//4)*ck.block_bytes(codec),dtype=np.uint8)
    for swap in ((0,1) if codec!=2 else (0,)):
        got=np.zeros(h*w*nc,np.uint8); lib.emu_decode4x4(codec,P(blocks),h,w,w//4,swap,P(got),w*nc)
        if not np.array_equal(got,ck.oracle_decode(codec,blocks,h,w,swap_rb=swap)): bad.append(("decode",codec,swap))
print("decode done",round(time.time()-t0,1),bad)
# downsample / pad on random blocks
for codec in (0,1,2):
    for (h,w) in ((512,512),(8,2048),(2048,8),(4,4),(4,64)):
        blocks=rng.integers(0,256,ck.nblocks(h)*ck.nblocks(w)*ck.block_bytes(codec),dtype=np.uint8)
        for st in ((0,1,2,3) if codec==2 else (2,)):
            want=ck.oracle_downsample(codec,blocks,h,w,strategy=st)
            got=np.zeros(want.size,np.uint8); s=lib.emu_downsample4x4(codec,st,P(blocks),h,w,P(got))
            if s!=0 or not np.array_equal(got,want): bad.append(("down",codec,h,w,st,s))
    for (ch,cw,ph,pw) in ((64,64,100,130),(8,8,8,64),(8,8,64,8),(128,256,129,257)):
        blocks=rng.integers(0,256,ck.nblocks(ch)*ck.nblocks(cw)*ck.block_bytes(codec),dtype=np.uint8)
        for st in ((2,3) if codec==2 else (2,)):
            want=ck.oracle_pad(codec,blocks,ch,cw,ph,pw,strategy=st)
            got=np.zeros(want.size,np.uint8); s=lib.emu_pad4x4(codec,st,P(blocks),ch,cw,ph,pw,P(got))
            if s!=0 or not np.array_equal(got,want): bad.append(("pad",codec,ch,cw,ph,pw,st,s))
blocks=rng.integers(0,256,8*500000,dtype=np.uint8); got=blocks.copy(); lib.emu_transcode_dxt1_to_etc1(P(got),got.size//8)
if not np.array_equal(got,ck.oracle_transcode(blocks)): bad.append(("transcode",))
print("blockops done",round(time.time()-t0,1),bad)
# PVRTC bigger images, more seeds
for n in (128,256):
    for kind in imagegen.KINDS:
        for seed in (1,2,3):
            img=imagegen.make(kind,n,n,4,seed=seed*7+n); flat=np.ascontiguousarray(img.ravel()); want=ck.oracle_pvrtc(flat,n,n)
            for parts in (1,5,8):
                out=np.zeros(n*n//4,np.uint8); s=lib.emu_pvrtc2(P(flat),n,n,parts,P(out))
                if s!=0 or not np.array_equal(out,want): bad.append(("pvrtc",n,kind,seed,parts,s))
print("pvrtc done",round(time.time()-t0,1),bad)
# ETC1 uniform random, 250K blocks per strategy
h,w=2000,2000
img=rng.integers(0,256,(h,w,3),dtype=np.uint8); flat=np.ascontiguousarray(img.ravel())
for st in (0,1,2,3):
    want=ck.oracle_etc1(st,flat,h,w); got=np.zeros(want.size,np.uint8)
    lib.emu_encode4x4(2,3,P(flat),h,w,w*3,h,w,0,st,P(got))
    if not np.array_equal(got,want): bad.append(("etc1",st))
print("etc1 done",round(time.time()-t0,1),bad)
# DXT all formats on random + smooth large
for fmt in (ck.RGB,ck.BGR,ck.RGBA,ck.BGRA):
    nc=ck.ncomp(fmt); codec=0 if nc==3 else 1; swap=1 if fmt in (ck.BGR,ck.BGRA) else 0
    for kind in ("random","smooth_noise","narrow","two_colour","alpha_extremes"):
        img=imagegen.make(kind,1024,1024,nc,seed=99); flat=np.ascontiguousarray(img.ravel()); want=ck.oracle_dxt(fmt,flat,1024,1024)
        for vote in (0,1):
            lib.emu_set_vote(vote); got=np.zeros(want.size,np.uint8)
            lib.emu_encode4x4(codec,nc,P(flat),1024,1024,1024*nc,1024,1024,swap,2,P(got))
            if not np.array_equal(got,want): bad.append(("dxt",fmt,kind,vote))
print("dxt done",round(time.time()-t0,1),bad)
print("TOTAL MISMATCHES",len(bad))
sys.exit(1 if bad else 0)
