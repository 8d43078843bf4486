"""Tiny on-disk datasets in the three formats the reference configs read (RESISC45 class folders,
DIOR COCO json, Potsdam tile/label PNG pairs), generated into a tmp dir by the tests."""
import json
import os

import cv2
import numpy as np

DIOR_CLASSES = ('airplane', 'airport', 'baseballfield', 'basketballcourt', 'bridge', 'chimney', 'dam',
                'Expressway-Service-area', 'Expressway-toll-station', 'golffield', 'groundtrackfield', 'harbor', 'overpass',
                'ship', 'stadium', 'storagetank', 'tenniscourt', 'trainstation', 'vehicle', 'windmill')


def make_resisc(root, classes=('airport', 'beach', 'forest'), per_class=4, size=72, seed=0):
    rng = np.random.default_rng(seed)
    for split in ('train', 'val'):
        for ci, c in enumerate(classes):
            d = os.path.join(root, split, c)
            os.makedirs(d, exist_ok=True)
            for i in range(per_class):
                img = rng.integers(0, 255, (size, size, 3), dtype=np.uint8)
                img[..., ci % 3] = 255 - 40 * ci                      # a learnable class cue
                cv2.imwrite(os.path.join(d, '%s_%03d.jpg' % (c, i)), img)
    return root


def make_dior(root, n_img=5, seed=0, hw=(80, 100)):
    """images with bright rectangles exactly where the annotated boxes are."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, 'JPEGImages-trainval'), exist_ok=True)
    os.makedirs(os.path.join(root, 'coco_ann'), exist_ok=True)
    images, anns = [], []
    aid = 1
    for i in range(n_img):
        h, w = hw if i % 2 == 0 else hw[::-1]                         # both aspect-ratio groups
        img = np.full((h, w, 3), 30, dtype=np.uint8)
        name = '%05d.png' % i
        images.append(dict(id=100 + i, file_name=name, width=w, height=h))
        for _ in range(0 if i == n_img - 1 else 2):                   # the last image has no gt (filtered in training)
            bw, bh = int(rng.integers(12, 30)), int(rng.integers(12, 30))
            x, y = int(rng.integers(0, w - bw)), int(rng.integers(0, h - bh))
            img[y:y + bh, x:x + bw] = 230
            anns.append(dict(id=aid, image_id=100 + i, category_id=int(rng.integers(1, 21)), bbox=[x, y, bw, bh],
                             area=bw * bh, iscrowd=0))
            aid += 1
        cv2.imwrite(os.path.join(root, 'JPEGImages-trainval', name), img)
    cats = [dict(id=k + 1, name=n) for k, n in enumerate(DIOR_CLASSES)]
    for split in ('train', 'val'):
        with open(os.path.join(root, 'coco_ann', 'DIOR_%s_coco.json' % split), 'w') as f:
            json.dump(dict(images=images, annotations=anns, categories=cats), f)
    return root


def make_potsdam(root, n=3, size=96, seed=0):
    """label PNG values 1..6 (0 never occurs, like the converted Potsdam labels); image = 40 * label on every channel."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    for split in ('train', 'val'):
        os.makedirs(os.path.join(root, 'img_IRRG', split), exist_ok=True)
        os.makedirs(os.path.join(root, 'ann_all', split), exist_ok=True)
        for i in range(n):
            lab = np.zeros((size, size), dtype=np.uint8)
            for by in range(0, size, 24):
                for bx in range(0, size, 24):
                    lab[by:by + 24, bx:bx + 24] = rng.integers(1, 7)
            img = np.repeat((lab * 40)[..., None], 3, -1).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(root, 'img_IRRG', split, 'tile_%d.png' % i))
            Image.fromarray(lab).save(os.path.join(root, 'ann_all', split, 'tile_%d.png' % i))
    return root
