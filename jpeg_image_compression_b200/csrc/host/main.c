/* main.c -- jpeg_compression_app <input_file_path> <output_file_path>
 * Same argv contract, messages and exit codes as natural_c/src/main.c:4-35
 * (exit 0 even if the save fails, "Save is sucesfull" without newline). */
#include <stdio.h>

#include "jpegb200.h"

int main(int argc, char *argv[])
{
    if (argc != 3) {
        fprintf(stderr, "Usage: %s <input_file_path> <output_file_path>\n", argv[0]);
        return 1;
    }
    printf("Starting processing...\n");
    printf("Input: %s\n", argv[1]);
    BMPImage *img = loadBMPImage(argv[1]);
    if (!img) {
        fprintf(stderr, "Error: Failed to load image from %s\n", argv[1]);
        return 1;
    }
    if (saveJPEGGrayscale(argv[2], img)) printf("Save is sucesfull");
    return 0;
}
