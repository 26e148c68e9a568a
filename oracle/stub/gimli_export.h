/* Stand-in for the cmake-generated export header (GenerateExportHeader). */
#ifndef GIMLI_EXPORT_H
#define GIMLI_EXPORT_H
#define DLLEXPORT
#define GIMLI_NO_EXPORT
#endif
