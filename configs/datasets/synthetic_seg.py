# Potsdam-shaped synthetic segmentation data (5 classes)
data = dict(samples_per_gpu=2, workers_per_gpu=0)
