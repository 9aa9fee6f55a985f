python -c "
import sys; sys.path.insert(0,'.')
import bench; open('/tmp/ref5m.fasta','wb').write(bench.synthetic_fasta())"
for q in 400x 2000x; do
time env RUSTYREAD_B200_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29562 -m rustyread_b200 -v simulate --reference /tmp/ref5m.fasta --quantity $q --seed 42 --error_model random --qscore_model random --output /dev/shm/rr_par.fastq --number_base_store 2G 2>&1 | grep -v "OMP_NUM\|\*\*\*\*"
ls -la /dev/shm/rr_par.fastq; rm -f /dev/shm/rr_par.fastq
done
time env RUSTYREAD_B200_TIMING=1 python -m rustyread_b200 -v simulate --reference /tmp/ref5m.fasta --quantity 2000x --seed 42 --error_model random --qscore_model random --output /dev/shm/rr_one.fastq --number_base_store 2G 2>&1 | tail -4
rm -f /dev/shm/rr_one.fastq
