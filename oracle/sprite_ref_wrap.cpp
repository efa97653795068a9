// TEST INFRASTRUCTURE ONLY - C-linkage door to the reference's own get_rg2s_cpp.
// Compiled together with the UNMODIFIED reference source where it lies
// (/root/reference/igm/cython_compiled/cpp_sprite_assignment.cpp) by oracle/Makefile
// into oracle/_ref/libsprite_ref.so; nothing of the reference is copied into this
// repository.  The declaration below restates igm/cython_compiled/cpp_sprite_assignment.h.
void get_rg2s_cpp(float* crds, int n_struct, int n_bead, int n_regions, int* copies_num,
                  float* rg2s, int* copy_idxs, int* min_struct);

extern "C" void ref_get_rg2s(float* crds, int n_struct, int n_bead, int n_regions, int* copies_num,
                             float* rg2s, int* copy_idxs, int* min_struct) {
    get_rg2s_cpp(crds, n_struct, n_bead, n_regions, copies_num, rg2s, copy_idxs, min_struct);
}
