# RESISC-45-shaped synthetic classification data (45 classes)
data = dict(samples_per_gpu=16, workers_per_gpu=0)
