// stubs/mfem.hpp -- DECLARATIONS ONLY of the subset of MFEM the -DPARELAGMC_B200_WITH_PARELAG branch of the host layer
// uses (compile-only check, see stubs/mpi.h).  Signatures as MFEM 3.x documents them; nothing here is ever linked.
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include "mpi.h"

namespace mfem {
template <class T> class Array {
public:
    Array();
    explicit Array(int n);
    Array(T *data, int n);
    void SetSize(int n);
    int Size() const;
    T *GetData();
    const T *GetData() const;
    T &operator[](int i);
    const T &operator[](int i) const;
    Array<T> &operator=(const T &v);
    T Max() const;
    int Find(const T &v) const;
};
class Vector {
public:
    Vector();
    explicit Vector(int n);
    Vector(double *data, int n);
    void SetSize(int n);
    int Size() const;
    double *GetData();
    const double *GetData() const;
    double &operator()(int i);
    const double &operator()(int i) const;
    double &operator[](int i);
    const double &operator[](int i) const;
    Vector &operator=(double v);
    Vector &operator=(const Vector &v);
    double Max() const;
    double Min() const;
};
class DenseMatrix {
public:
    int Height() const;
    int Width() const;
    double *Data();
};
class SparseMatrix {
public:
    int Height() const;
    int Width() const;
    int Size() const;
    const int *GetI() const;
    const int *GetJ() const;
    const double *GetData() const;
    int NumNonZeroElems() const;
    void GetDiag(Vector &d) const;
    void EliminateRowCol(int rc, int diag_policy = 0);
    void EliminateCols(const Array<int> &cols, const Vector *x = nullptr, Vector *b = nullptr);
    void ScaleRows(const Vector &s);
    void Mult(const Vector &x, Vector &y) const;
    void MultTranspose(const Vector &x, Vector &y) const;
};
SparseMatrix *Mult(const SparseMatrix &A, const SparseMatrix &B);
SparseMatrix *Transpose(const SparseMatrix &A);
class HypreParMatrix {
public:
    int Height() const;
    int Width() const;
    void Mult(const Vector &x, Vector &y) const;
    void MultTranspose(const Vector &x, Vector &y) const;
};
class Coefficient { public: virtual ~Coefficient(); };
class ConstantCoefficient : public Coefficient { public: explicit ConstantCoefficient(double c = 1.0); };
class VectorCoefficient { public: virtual ~VectorCoefficient(); };
class BilinearFormIntegrator { public: virtual ~BilinearFormIntegrator(); };
class LinearFormIntegrator { public: virtual ~LinearFormIntegrator(); };
class VectorFEMassIntegrator : public BilinearFormIntegrator { public: explicit VectorFEMassIntegrator(Coefficient &q); };
class FiniteElementCollection { public: virtual ~FiniteElementCollection(); };
class L2_FECollection : public FiniteElementCollection { public: L2_FECollection(int p, int dim); };
class Mesh {
public:
    int Dimension() const;
    int GetNE() const;
    Array<int> bdr_attributes;
    void UniformRefinement();
    virtual void Print(std::ostream &os = std::cout) const;
};
class ParMesh : public Mesh {
public:
    ParMesh(MPI_Comm comm, Mesh &mesh);
    MPI_Comm GetComm() const;
};
class FiniteElementSpace {
public:
    FiniteElementSpace(Mesh *m, const FiniteElementCollection *fec);
    int GetNDofs() const;
};
class GridFunction : public Vector {
public:
    GridFunction();
    explicit GridFunction(FiniteElementSpace *f);
    void MakeRef(FiniteElementSpace *f, Vector &v, int v_offset);
    void ProjectBdrCoefficientNormal(VectorCoefficient &vcoeff, Array<int> &bdr_attr);
    double ComputeL2Error(Coefficient &exsol) const;
    virtual void Save(std::ostream &out) const;
};
}  // namespace mfem

// ---- linear forms and integrators used by DarcySolver's host-once set-up (src/DarcySolver.cpp:246-414) ----
namespace mfem {
class VectorConstantCoefficient : public VectorCoefficient { public: explicit VectorConstantCoefficient(const Vector &v); };
class RestrictedCoefficient : public Coefficient { public: RestrictedCoefficient(Coefficient &c, Array<int> &attr); };
class DomainLFIntegrator : public LinearFormIntegrator { public: explicit DomainLFIntegrator(Coefficient &q); };
class VectorFEDomainLFIntegrator : public LinearFormIntegrator { public: explicit VectorFEDomainLFIntegrator(VectorCoefficient &f); };
class VectorFEBoundaryFluxLFIntegrator : public LinearFormIntegrator { public: explicit VectorFEBoundaryFluxLFIntegrator(Coefficient &f); };
class LinearForm : public Vector {
public:
    LinearForm();
    void AddDomainIntegrator(LinearFormIntegrator *lfi);
    void AddBoundaryIntegrator(LinearFormIntegrator *lfi);
    void Update(FiniteElementSpace *f, Vector &v, int v_offset);
    void Assemble();
};
class BlockMatrix {
public:
    explicit BlockMatrix(const Array<int> &offsets);
    void SetBlock(int i, int j, SparseMatrix *m);
    void MultTranspose(const Vector &x, Vector &y) const;
    void Mult(const Vector &x, Vector &y) const;
    int owns_blocks;
};
}  // namespace mfem
