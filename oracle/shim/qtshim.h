// TEST INFRASTRUCTURE ONLY (oracle shim) -- never linked into the product.
// Stand-ins for the handful of Qt4 types the reference's collision sources mention
// (QString/QTextStream/QPixmap/QMessageBox via include/graspit/mytools.h; QMutex and
// QThreadStorage via include/graspit/Collision/collisionInterface.h).  Qt4 is absent
// from this image.
#ifndef B2C_ORACLE_QTSHIM_H
#define B2C_ORACLE_QTSHIM_H
#include <string>
#include <list>
#include <cstdio>

class QString {
    std::string s_;
  public:
    QString() {}
    QString(const char *c) : s_(c ? c : "") {}
    QString(const std::string &s) : s_(s) {}
    std::string toStdString() const { return s_; }
    const char *latin1() const { return s_.c_str(); }
    bool isNull() const { return s_.empty(); }
    bool operator==(const QString &o) const { return s_ == o.s_; }
    bool operator!=(const QString &o) const { return s_ != o.s_; }
    QString operator+(const QString &o) const { return QString(s_ + o.s_); }
    QString &setNum(int n) { s_ = std::to_string(n); return *this; }
};
class QTextStream {};
class QPixmap {};
namespace Qt { enum { NoButton = 0 }; }
class QMessageBox {
  public:
    enum { Ok = 1 };
    static int warning(void *, const char *, const QString &m, int, int, int) {
      std::fprintf(stderr, "%s\n", m.latin1()); return 0;
    }
};
class QMutex {
  public:
    void lock() {}
    void unlock() {}
};
// one slot per object; the oracle driver is single-threaded per world instance
template <class T> class QThreadStorage {
    T data_;
    bool set_;
  public:
    QThreadStorage() : data_(), set_(false) {}
    void setLocalData(T d) { data_ = d; set_ = true; }
    T localData() const { return set_ ? data_ : T(); }
    bool hasLocalData() const { return set_; }
};
#endif
