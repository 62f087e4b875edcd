def imread(path, **kwargs):
    import imageio
    return imageio.imread(path)


def imsave(path, arr, **kwargs):
    import imageio
    imageio.imwrite(path, arr)
