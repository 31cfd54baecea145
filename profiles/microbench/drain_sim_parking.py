import numpy as np, heapq
rng=np.random.default_rng(1)
N=1_250_000; LANES=113664; W=LANES//32
bins=[(7,.362),(8,.486),(9,.115),(10,.018),(11,.010),(12,.006),(13,.003)]
p=np.array([b[1] for b in bins]); p/=p.sum()
b=rng.choice([x[0] for x in bins],size=N,p=p)
pops=(2.0**(b+rng.random(N))).astype(np.int64)
ideal=pops.sum()/LANES
def fast_pass(order):
    q=pops[order]
    heap=[(0.0,l) for l in range(LANES)]; heapq.heapify(heap)
    for w in q:
        tt,l=heapq.heappop(heap); heapq.heappush(heap,(tt+w,l))
    fin=np.zeros(LANES)
    for tt,l in heap: fin[l]=tt
    dry=min(tt for tt,_ in heap)
    return fin, dry
def suspend_round(fin_w, t0, T):
    """fin_w: [warps,32] finish times of the reads in flight (absolute), t0 = time after which suspension is allowed.
    returns slot_time used by these warps after t0, and remaining work of suspended reads"""
    s=np.sort(fin_w,axis=1)            # ascending finish
    # warp runs until only T lanes remain active: i.e. until the (32-T)-th lane finishes (index 31-T), but not before t0
    if T>0:
        stop=np.maximum(s[:,31-T],t0)
    else:
        stop=np.maximum(s[:,31],t0)
    slot=(stop-t0).clip(min=0).sum()*32
    rem=(fin_w-stop[:,None]).clip(min=0)
    rem=rem[rem>0]
    return slot, rem, stop
for label,order in (("random",rng.permutation(N)),("noisy class",np.argsort(-(b+rng.normal(0,1.0,N)),kind='stable'))):
    fin,dry=fast_pass(order)
    base_slot=dry*LANES
    for T in (0,8,16,24):
        slot,rem,stop=suspend_round(fin.reshape(W,32),dry,T)
        total=base_slot+slot; rounds=0; lat=stop.max()
        while T>0 and len(rem)>0:
            rounds+=1
            n=len(rem); nw=(n+31)//32
            pad=np.zeros(nw*32); pad[:n]=rng.permutation(rem)
            Tn=T if n>4096 else 0
            slot,rem2,stop=suspend_round(pad.reshape(nw,32),0.0,Tn)
            total+=slot; lat+=stop.max(); rem=rem2
            if Tn==0: break
        print(f"{label:12s} T={T:2d} slot-time/ideal {total/(ideal*LANES):.3f} rounds {rounds} batch latency {lat:.0f} (ideal {ideal:.0f})")
print("--- park once (T at fast pass), then one dense resume launch run to completion; and two rounds")
for label,order in (("random",rng.permutation(N)),("noisy class",np.argsort(-(b+rng.normal(0,1.0,N)),kind='stable'))):
    fin,dry=fast_pass(order)
    base_slot=dry*LANES
    for T in (8,16,24):
        for nr in (1,2,3):
            slot,rem,stop=suspend_round(fin.reshape(W,32),dry,T)
            total=base_slot+slot; lat=stop.max()
            for rr in range(nr):
                if len(rem)==0: break
                n=len(rem); nw=(n+31)//32
                pad=np.zeros(nw*32); pad[:n]=rng.permutation(rem)
                Tn=T if rr<nr-1 else 0
                slot,rem,stop=suspend_round(pad.reshape(nw,32),0.0,Tn)
                total+=slot; lat+=stop.max()
            print(f"{label:12s} T={T:2d} rounds={nr} slot-time/ideal {total/(ideal*LANES):.3f} latency {lat:.0f}")
