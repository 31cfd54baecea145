import numpy as np, heapq, sys
rng=np.random.default_rng(1)
N=1_250_000; LANES=113664; W=LANES//32
bins=[(7,.362),(8,.486),(9,.115),(10,.018),(11,.010),(12,.006),(13,.003)]
p=np.array([b[1] for b in bins]); p/=p.sum()
b=rng.choice([x[0] for x in bins],size=N,p=p)
pops=(2.0**(b+rng.random(N))).astype(np.int64)
print("mean pops",pops.mean(),"max",pops.max(), "total/lanes", pops.sum()/LANES)
def run(order, label):
    # lanes grouped in warps; each lane takes next read from queue when done. time unit = iterations (constant tau)
    q=pops[order]
    t=np.zeros(LANES)            # lane finish times
    heap=[(0.0,l) for l in range(LANES)]
    heapq.heapify(heap)
    for w in q:
        tt,l=heapq.heappop(heap)
        heapq.heappush(heap,(tt+w,l))
    fin=np.zeros(LANES)
    for tt,l in heap: fin[l]=tt
    dry=min(tt for tt,_ in heap) # approx time the queue ran dry (first lane to find no work)
    warp_end=fin.reshape(W,32).max(axis=1)
    ideal=pops.sum()/LANES
    idle=(warp_end[:,None]-fin.reshape(W,32)).sum()
    print(f"{label:28s} ideal {ideal:7.0f} dry {dry:7.0f} makespan {fin.max():7.0f}  warp exit pct10/50/90 {np.percentile(warp_end,[10,50,90]).round()}  idle lane-iters/ideal total {idle/(ideal*LANES):.3f}")
run(rng.permutation(N),"random")
run(np.argsort(-pops,kind='stable'),"exact LPT")
# noisy class: class = log2 bin with noise
noisy=np.argsort(-(b+rng.normal(0,1.0,N)),kind='stable'); run(noisy,"noisy class (sd 1 bin)")
cls=np.argsort(-b,kind='stable'); run(cls,"exact bin class")
