import time, sys
sys.path.insert(0, "/root/repo")
import rustyread_b200 as rr
t=time.time(); c=rr.Context(0); t1=time.time(); c.close(); t2=time.time()
c=rr.Context(0); t3=time.time()
print("context create %.1f ms, close %.1f ms, second create %.1f ms" % ((t1-t)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
