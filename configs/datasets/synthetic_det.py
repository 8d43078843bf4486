# DIOR-shaped synthetic detection data (20 classes, 8 boxes / image)
data = dict(samples_per_gpu=1, workers_per_gpu=0)
