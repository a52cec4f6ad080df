"""Drop-in for the reference's convert.py (PNG sequence -> GIF).  Host-side I/O utility,
not part of the hot path; same signature as convert.py:4 of the reference."""
import numpy as np


def convert(directory, filename, start, stop, skip, outname):
    import imageio as iio          # the reference imports imageio at module import time
    images = []
    for i in np.arange(start, stop, skip):
        file = directory + '/' + filename + '_' + str(i) + '.png'
        images.append(iio.imread(file))
    iio.mimsave(outname, images, duration=0.2)


if __name__ == '__main__':
    convert('plots', 'summary', 0, 1000, 10, 'movie_all.gif')
